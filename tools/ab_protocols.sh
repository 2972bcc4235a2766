#!/bin/bash
# The three exchange protocols under the multi-GPU bench, on one box:
#   gpurun --gpus N -- bash tools/ab_protocols.sh N [rounds]
N="$1"; R="${2:-2}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/abp_n${N}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/abp_n${N}_pytest_multi.log
tail -3 gpurun_out/abp_n${N}_pytest_multi.log
for round in $(seq 1 $R); do
for proto in 3 2 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 300 --warmup 20 --strong 0 --verify 0 --protocol $proto > gpurun_out/abp_n${N}_p${proto}_${round}.json 2> gpurun_out/abp_n${N}_p${proto}_${round}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abp_n${N}_p${proto}_${round}.json"))
    print("N=${N} protocol ${proto} r${round}: step %.2f us  sustained %.2f  bracketed %.2f  e2e %.1f  phases %s" % (
        1e3 * d["ms_per_step"], 1e3 * d["sustained"]["ms_per_step"], d["roofline"]["avg_launch_us"],
        1e3 * d["e2e"]["ms_per_step"], json.dumps(d.get("phases_rank0_us"))))
except Exception as e:
    print("protocol ${proto} r${round}: ERR", e)
PY
done
done
