#!/bin/bash
# N=8 weak-scaling bench only (8x GPU-minutes: keep it as short as possible)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 500 --warmup 10 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err; echo "rc=$?"; cat gpurun_out/scale_n8.json | cut -c1-300
