#!/bin/bash
# A/B of library builds under the multi-GPU bench, on one box:
#   gpurun --gpus N -- bash tools/ab_multi.sh N libA.so libB.so ...
N="$1"; shift
mkdir -p gpurun_out
for round in 1 2; do
for lib in "$@"; do
  tag=$(basename $lib .so)
  CFMM_B200_LIB=$PWD/$lib timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 300 --warmup 20 --strong 0 --verify 0 > gpurun_out/abm_${tag}_${round}.json 2> gpurun_out/abm_${tag}_${round}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/abm_${tag}_${round}.json"))
    print("${tag} r${round}: step %.2f us  sustained %.2f  bracketed %.2f  e2e %.1f  phases %s" % (
        1e3 * d["ms_per_step"], 1e3 * d["sustained"]["ms_per_step"], d["roofline"]["avg_launch_us"],
        1e3 * d["e2e"]["ms_per_step"], json.dumps(d.get("phases_rank0_us"))))
except Exception as e:
    print("${tag} r${round}: ERR", e)
PY
done
done
