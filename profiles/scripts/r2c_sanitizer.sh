#!/bin/bash
# compute-sanitizer over the kernels added in the last session of round 2: giant-cell sort / collide (two cells of 70 000 parcels),
# mesh-wide locate, cell-order translation, the unrolled index kernels of the sort
cd /root/repo
mkdir -p gpurun_out
echo "=== memcheck ==="
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_chemistry.py tests/test_gpu_edge_cases.py tests/test_gpu_cell_order.py -q -m gpu -x -k "not shipped_series and not equilibrium" 2>&1 | tail -25 | tee gpurun_out/r02c_memcheck.log
echo "=== racecheck (giant-cell kernels, sort) ==="
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_chemistry.py -q -m gpu -x -k "two-cells-of-70000" 2>&1 | tail -40 | tee gpurun_out/r02c_racecheck.log
echo "=== synccheck ==="
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_chemistry.py tests/test_gpu_edge_cases.py -q -m gpu -x -k "two-cells-of-70000 or thousands or located" 2>&1 | tail -15 | tee gpurun_out/r02c_synccheck.log
