#!/bin/bash
# N=8 pass: multi-GPU tests over all ranks, then the scaling ladder the driver runs (N=1,2,4,8)
set -x
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -8 gpurun_out/pytest_multi.log
for N in 1 2 4 8; do
  if [ $N -gt $NG ]; then continue; fi
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 1000 --warmup 10 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 1000 --warmup 10 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$?"; cat gpurun_out/scale_n$N.json; tail -2 gpurun_out/scale_n$N.err
done
if [ $NG -ge 8 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 1000 --warmup 10 --exchange nccl > gpurun_out/scale_n8_nccl.json 2> gpurun_out/scale_n8_nccl.err; cat gpurun_out/scale_n8_nccl.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 1000 --warmup 10 --scaling strong > gpurun_out/scale_n8_strong.json 2> gpurun_out/scale_n8_strong.err; cat gpurun_out/scale_n8_strong.json
fi
