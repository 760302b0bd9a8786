#!/bin/bash
# A/B of the move-kernel instance without wall functions (periodic box, no patch carries a wall model) against the instance with them,
# same box: DSMCB200_MOVE_WALLS=1 forces the latter.  ms per step / per kernel.
cd /root/repo
run() { # <cells> <steps> <walls>
  DSMCB200_MOVE_WALLS=$3 python bench.py --cells $1 --gas air5 --steps $2 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; print('cells $1 walls $3', round(d['ms_per_step'],3), {a:round(b,3) for a,b in k.items() if a in ('move','collide','sample','gather')})"
}
run 100 6 1; run 100 6 0; run 100 6 1; run 100 6 0
run 200 8 1; run 200 8 0
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
