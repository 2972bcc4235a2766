#!/bin/bash
# ncu summaries of the secondary kernels (GeometricMean, UniV3) for profiles/
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:sweep_kernel -s 4 -c 1 -o gpurun_out/prof_r1_geomean python bench.py --workload config3_1M_mixed_10k_tokens --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_geomean.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:sweep_kernel -s 4 -c 1 -o gpurun_out/prof_r1_univ3 python bench.py --workload config4_500k_univ3_5k_tokens --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_univ3.log 2>&1
for w in config3_1M_mixed_10k_tokens config4_500k_univ3_5k_tokens config2_100k_product_1k_tokens; do timeout 300 python bench.py --workload $w --steps 1000 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_$w.json; cat gpurun_out/bench_$w.json | cut -c1-400; done
