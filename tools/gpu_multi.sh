#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -12 gpurun_out/pytest_multi.log
for ts in 0 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ts bench.py --gpus 2 --steps 1000 --warmup 10 --two-shot $ts > gpurun_out/bench_n2_ts$ts.json 2> gpurun_out/bench_n2_ts$ts.err; echo "rc=$?"
  cat gpurun_out/bench_n2_ts$ts.json
done
