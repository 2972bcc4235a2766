#!/bin/bash
# N=2 pass: multi-GPU tests, 2-GPU bench (peer and nccl exchange, weak and strong), plus 1-GPU variant sweep
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
for ex in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 10 --exchange $ex > gpurun_out/bench_n2_$ex.json 2> gpurun_out/bench_n2_$ex.err; echo "rc=$?" >> gpurun_out/bench_n2_$ex.err
  cat gpurun_out/bench_n2_$ex.json; tail -3 gpurun_out/bench_n2_$ex.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1000 --warmup 10 --scaling strong > gpurun_out/bench_n2_strong.json 2> gpurun_out/bench_n2_strong.err
cat gpurun_out/bench_n2_strong.json; tail -3 gpurun_out/bench_n2_strong.err
timeout 300 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json

