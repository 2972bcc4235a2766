#!/bin/bash
# One single-GPU call: parity tests, bench lines of every BASELINE config, phase trace, ncu captures.
#   gpurun --timeout 2400 -- bash tools/gpu_round2.sh [tag]
tag="${1:-r2}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${tag}_clocks.csv &
SMI=$!
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log; tail -3 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench_n1.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n1_k20.json 2>> gpurun_out/${tag}_bench_n1.err
for wl in config2_100k_product_1k_tokens config3_1M_mixed_10k_tokens config4_500k_univ3_5k_tokens; do
  timeout 600 python bench.py --workload $wl --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/${tag}_bench_$wl.json 2>> gpurun_out/${tag}_bench_n1.err
  cat gpurun_out/${tag}_bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_n1.err; cat gpurun_out/${tag}_bench_reference.json
kill $SMI
python tools/trace_phases.py --json gpurun_out/${tag}_trace_10M.json > gpurun_out/${tag}_trace.log 2>&1
python tools/trace_phases.py --m 1250000 --json gpurun_out/${tag}_trace_1250k.json >> gpurun_out/${tag}_trace.log 2>&1
python tools/trace_phases.py --opt balance=0 --json gpurun_out/${tag}_trace_10M_nobalance.json >> gpurun_out/${tag}_trace.log 2>&1
cat gpurun_out/${tag}_trace.log
# ncu: launch list of the bench command, then one full capture per kernel family
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:product_sweep_tma -s 8 -c 2 -o gpurun_out/${tag}_prof_product python tools/ab_kernel.py --iters 4 --rounds 1 > gpurun_out/${tag}_ncu_product.log 2>&1
ncu --set full --clock-control none -k regex:product_sweep_tma -s 6 -c 1 -o gpurun_out/${tag}_prof_geomean python bench.py --workload config3_1M_mixed_10k_tokens --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_ncu_geomean.log 2>&1
ncu --set full --clock-control none -k regex:sweep_kernel -s 6 -c 1 -o gpurun_out/${tag}_prof_univ3 python bench.py --workload config4_500k_univ3_5k_tokens --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_ncu_univ3.log 2>&1
ls -la gpurun_out | tail -30
