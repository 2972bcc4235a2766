#!/bin/bash
# One single-GPU call: parity tests, smoke, bench lines of every BASELINE config, A/B runs, phase
# traces, ncu captures.   gpurun --timeout 2400 -- bash tools/gpu_round2.sh [tag]
tag="${1:-r2}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/${tag}_clocks.csv &
SMI=$!
timeout 1200 python -m pytest tests -x -q -m gpu > $O/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${tag}_pytest_gpu.log
tail -4 $O/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${tag}_smoke.log; tail -3 $O/${tag}_smoke.log
CFMM_TIMING=1 timeout 900 python bench.py > $O/${tag}_bench_n1.json 2> $O/${tag}_bench_n1.err; echo "bench rc=$?"; cat $O/${tag}_bench_n1.json
grep "\[cfmm\]" $O/${tag}_bench_n1.err > $O/${tag}_finalize_timing.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${tag}_bench_n1_k20.json 2>> $O/${tag}_bench_n1.err
for wl in config2_100k_product_1k_tokens config3_1M_mixed_10k_tokens config4_500k_univ3_5k_tokens; do
  timeout 600 python bench.py --workload $wl --steps 500 --warmup 20 --no-cpu-baseline > $O/${tag}_bench_$wl.json 2>> $O/${tag}_bench_n1.err
  cat $O/${tag}_bench_$wl.json
done
timeout 600 python bench.py --workload config4_500k_univ3_5k_tokens --steps 500 --warmup 20 --no-cpu-baseline --opt grid_waves=1 > $O/${tag}_bench_config4_grid_waves1.json 2>> $O/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${tag}_bench_reference.json 2>> $O/${tag}_bench_n1.err; cat $O/${tag}_bench_reference.json
kill $SMI
# A/B on this box: the round-1 build against this one (headline kernel), option ablations, GeometricMean kernels
python tools/ab_kernel.py --lib r1=gpurun_scratch/libcfmm_r1.so --lib new=cfmmrouter.jl_b200/libcfmm_b200.so --out $O/${tag}_ab_r1_vs_new.json > $O/${tag}_ab.log 2>&1
python tools/ab_kernel.py --opt "" --opt "compact_stream=0" --opt "balance=0" --opt "psi_fixed_point=0" --opt "gradient_math=0" --out $O/${tag}_ab_options.json >> $O/${tag}_ab.log 2>&1
python tools/ab_kernel.py --type geomean --m 5000000 --iters 100 --rounds 2 --opt "geomean_tma=1" --opt "geomean_tma=0" --out $O/${tag}_ab_geomean_5M.json >> $O/${tag}_ab.log 2>&1
python tools/ab_kernel.py --type geomean --m 500000 --n 10000 --iters 100 --rounds 2 --flush --opt "geomean_tma=1" --opt "geomean_tma=0" --out $O/${tag}_ab_geomean_500k_flushed.json >> $O/${tag}_ab.log 2>&1
grep median $O/${tag}_ab.log
python tools/e2e_breakdown.py > $O/${tag}_e2e_breakdown.json 2> $O/${tag}_e2e_breakdown.err; cat $O/${tag}_e2e_breakdown.json
python tools/trace_phases.py --json $O/${tag}_trace_10M.json > $O/${tag}_trace.log 2>&1
python tools/trace_phases.py --m 1250000 --json $O/${tag}_trace_1250k.json >> $O/${tag}_trace.log 2>&1
python tools/trace_phases.py --opt balance=0 --json $O/${tag}_trace_10M_nobalance.json >> $O/${tag}_trace.log 2>&1
cat $O/${tag}_trace.log
# ncu: launch list of the bench command, then one full capture per kernel family
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $O/${tag}_launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:product_sweep_tma -s 20 -c 2 -o $O/${tag}_prof_product python tools/ab_kernel.py --iters 4 --rounds 1 > $O/${tag}_ncu_product.log 2>&1
ncu --set full --clock-control none -k regex:product_sweep_tma -s 12 -c 2 -o $O/${tag}_prof_config3 python bench.py --workload config3_1M_mixed_10k_tokens --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/${tag}_ncu_config3.log 2>&1
ncu --set full --clock-control none -k regex:product_sweep_tma -s 12 -c 1 -o $O/${tag}_prof_geomean_5M python tools/ab_kernel.py --type geomean --m 5000000 --iters 4 --rounds 1 > $O/${tag}_ncu_geomean_5M.log 2>&1
ncu --set full --clock-control none -k regex:sweep_kernel -s 6 -c 1 -o $O/${tag}_prof_univ3 python bench.py --workload config4_500k_univ3_5k_tokens --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/${tag}_ncu_univ3.log 2>&1
ls -la $O | grep ${tag}_ | tail -50
