#!/bin/bash
# Lean multi-GPU call for large N (charged N x): gpurun --gpus N -- bash tools/gpu_multi8.sh N [tag]
N="${1:-8}"; tag="${2:-r2}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_n${N}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/${tag}_n${N}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_n${N}_pytest_multi.log
tail -3 gpurun_out/${tag}_n${N}_pytest_multi.log
run() { # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/${tag}_n${N}_${name}.json 2> gpurun_out/${tag}_n${N}_${name}.err
  echo "$name rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${tag}_n${N}_${name}.err | tail -3 | cut -c1-300
}
run bench --steps 300 --warmup 20
run bench_k20 --steps 20 --warmup 5          # what the driver runs
if [ "${3:-}" = "protocols" ]; then           # third argument: also compare the exchange protocols on this box
  run bench_direct --steps 300 --warmup 20 --strong 0 --verify 0 --protocol 3
  run bench_ll_twoshot --steps 300 --warmup 20 --strong 0 --verify 0 --protocol 2
  run bench_ll_oneshot --steps 300 --warmup 20 --strong 0 --verify 0 --protocol 1
fi
python - <<PY
import json
for name in ("bench", "bench_k20", "bench_direct", "bench_ll_twoshot", "bench_ll_oneshot"):
    try:
        d = json.load(open("gpurun_out/${tag}_n${N}_%s.json" % name))
    except Exception as e:
        continue
    print(name, "value %.4g step %.2f us e2e %.1f us bracketed %.2f parity %s sustained %.2f" % (d["value"], 1e3 * d["ms_per_step"], 1e3 * d["e2e"]["ms_per_step"], d["roofline"]["avg_launch_us"], d.get("parity_checked"), 1e3 * d["sustained"]["ms_per_step"]))
    for k in ("strong", "phases_rank0_us", "parity"):
        if k in d: print("  ", k, json.dumps(d[k])[:1500])
PY
