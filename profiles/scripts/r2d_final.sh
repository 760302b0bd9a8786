#!/bin/bash
# evidence run of the last session of round 2 (1 GPU, ABI version 5 / dsmcCLLWallPatch in a move-kernel instance of its own):
# GPU tests, default bench, capsule bench (a case with walls and inflow on the instance without the CLL kernel)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2d_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2d_final_box.json 2> gpurun_out/r2d_final_box.err; tail -c 300 gpurun_out/r2d_final_box.err
timeout 240 python bench.py --workload capsule --steps 10 --warmup 20 --no-cpu-baseline > gpurun_out/r2d_final_capsule.json 2> gpurun_out/r2d_final_capsule.err; tail -c 300 gpurun_out/r2d_final_capsule.err
python -c "
import json
for f in ('box','capsule'):
    try:
        d=json.loads(open('gpurun_out/r2d_final_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d.get('ms_per_step'), d.get('kernel_ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f, 'FAILED', e)
"
