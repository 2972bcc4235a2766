#!/bin/bash
# One GPU call: parity tests, then the 1-GPU bench line.  Usage: gpurun -- bash tools/gpu_check.sh [pytest -k expr]
mkdir -p gpurun_out
K="${1:-}"
if [ -n "$K" ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu -k "$K" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
