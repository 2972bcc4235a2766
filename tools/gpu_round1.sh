#!/bin/bash
# First GPU pass: smoke, parity tests, bench, launch list, one full ncu capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json
timeout 600 python tools/explore.py > gpurun_out/explore.log 2>&1; echo "explore rc=$?" >> gpurun_out/explore.log
cat gpurun_out/explore.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 5 -c 2 -o gpurun_out/prof_r1_product python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
