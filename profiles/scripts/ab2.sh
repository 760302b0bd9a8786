#!/bin/bash
# usage: ab2.sh <cells> <variants...>  (variant "base" = in-tree library; others build_variants/<name>.so)
cells=$1; shift 1
for v in "$@"; do
  if [ "$v" == "base" ]; then unset DSMCB200_LIB; else export DSMCB200_LIB=/root/repo/build_variants/$v.so; fi
  python bench.py --cells $cells --gas air5 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('$v', round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items()})"
done
