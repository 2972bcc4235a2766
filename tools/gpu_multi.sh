#!/bin/bash
# Multi-GPU call: gpurun --gpus N --timeout 1800 -- bash tools/gpu_multi.sh N [tag]
N="${1:-2}"; tag="${2:-r2}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_n${N}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/${tag}_n${N}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_n${N}_pytest_multi.log
tail -6 gpurun_out/${tag}_n${N}_pytest_multi.log
run() { # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/${tag}_n${N}_${name}.json 2> gpurun_out/${tag}_n${N}_${name}.err
  echo "$name rc=$?"; tail -2 gpurun_out/${tag}_n${N}_${name}.err | cut -c1-300; cat gpurun_out/${tag}_n${N}_${name}.json
}
run bench --steps 500 --warmup 20
run bench_k20 --steps 20 --warmup 3 --strong 0 --verify 0
run bench_coop --steps 200 --warmup 20 --strong 0 --verify 0 --opt coop_launch=1
run bench_nccl --steps 300 --warmup 10 --strong 0 --verify 0 --exchange nccl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${tag}_n${N}_reference.json 2>> gpurun_out/${tag}_n${N}_bench.err
cat gpurun_out/${tag}_n${N}_reference.json
