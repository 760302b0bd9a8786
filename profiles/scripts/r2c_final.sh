#!/bin/bash
# end-of-round evidence run of the last session of round 2 (1 GPU): GPU tests, default bench (e2e + CPU baseline), C1, wedge, cylinder, capsule,
# reference arm, ncu launch list + DRAM traffic + full capture of the final build
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2c_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2c_final_box.json 2> gpurun_out/r2c_final_box.err; tail -c 300 gpurun_out/r2c_final_box.err
timeout 600 python bench.py --workload c1 --steps 50 --warmup 10 > gpurun_out/r2c_final_c1.json 2> gpurun_out/r2c_final_c1.err; tail -c 300 gpurun_out/r2c_final_c1.err
timeout 600 python bench.py --workload wedge --steps 10 --warmup 20 > gpurun_out/r2c_final_wedge.json 2> gpurun_out/r2c_final_wedge.err; tail -c 300 gpurun_out/r2c_final_wedge.err
timeout 400 python bench.py --workload cylinder --steps 20 --warmup 5 > gpurun_out/r2c_final_cyl.json 2> gpurun_out/r2c_final_cyl.err; tail -c 300 gpurun_out/r2c_final_cyl.err
timeout 600 python bench.py --workload capsule --steps 10 --warmup 20 > gpurun_out/r2c_final_capsule.json 2> gpurun_out/r2c_final_capsule.err; tail -c 300 gpurun_out/r2c_final_capsule.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2c_final_ref.json 2> gpurun_out/r2c_final_ref.err; tail -c 300 gpurun_out/r2c_final_ref.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_air5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"moveKernel|gatherKernel|collideLaneKernel|sampleKernel" --launch-skip 12 --launch-count 4 --csv --log-file gpurun_out/r02c_traffic_box_air5.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"moveKernel|gatherKernel|collideLaneKernel|sampleKernel" --launch-skip 76 --launch-count 8 --csv --log-file gpurun_out/r02c_traffic_capsule.csv python bench.py --workload capsule --steps 1 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/ncu_traffic_c.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"moveKernel|gatherKernel|collideLaneKernel|sampleKernel" --launch-skip 12 --launch-count 4 -o gpurun_out/r02c_final_c100 python bench.py --cells 100 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
python profiles/scripts/hb_gpu.py 2>&1 | tail -2 | tee gpurun_out/r2c_heatbath.log
python -c "
import json
for f in ('box','c1','wedge','cyl','capsule','ref'):
    try:
        d=json.loads(open('gpurun_out/r2c_final_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d.get('ms_per_step'), d.get('kernel_ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('cpu_baseline'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f, 'FAILED', e)
"
