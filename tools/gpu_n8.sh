#!/bin/bash
# N=8 pass (8x GPU-minutes: keep it short): multi-GPU tests over all ranks, weak + strong bench at N=8
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi8.log
tail -8 gpurun_out/pytest_multi8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 500 --warmup 10 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err; echo "rc=$?"; cat gpurun_out/scale_n8.json; tail -2 gpurun_out/scale_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 500 --warmup 10 --scaling strong > gpurun_out/scale_n8_strong.json 2> gpurun_out/scale_n8_strong.err; cat gpurun_out/scale_n8_strong.json
