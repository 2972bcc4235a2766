#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/explore.py > gpurun_out/explore.log 2>&1; echo "explore rc=$?" >> gpurun_out/explore.log
cat gpurun_out/explore.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:product_sweep_tma -s 5 -c 1 -o gpurun_out/prof_r1_tma python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
