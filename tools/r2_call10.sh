set -x
mkdir -p gpurun_out
MVTM_RING=1 timeout 900 python tools/ab.py build_ab/libmvtm_base.so build_ab/libmvtm_v1.so build_ab/libmvtm_v5.so build_ab/libmvtm_v6.so acm_2v:200000 pubmed_3v:60000 lda_100k > gpurun_out/r2_ab_variants2.log 2>&1; cat gpurun_out/r2_ab_variants2.log
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_v6.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_v6.log
timeout 600 python tools/tail_exp.py 250000 > gpurun_out/r2_tail_exp.log 2>&1; cat gpurun_out/r2_tail_exp.log | tail -4
for wl in lda_100k uniform_k1000; do
timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${wl}_1gpu.json 2> gpurun_out/r2_bench_${wl}_1gpu.err; echo "$wl rc=$?"; tail -n 1 gpurun_out/r2_bench_${wl}_1gpu.json | cut -c1-250
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 8 -c 1 -o gpurun_out/prof_sweep_r2_uniform -f python tools/run_one.py uniform_k1000 10 200000 > gpurun_out/ncu_full_uniform.log 2>&1; tail -2 gpurun_out/ncu_full_uniform.log
timeout 600 python tools/steady_state.py acm_2v 100000 400 > gpurun_out/r2_steady_acm.log 2>&1; cat gpurun_out/r2_steady_acm.log | tail -22
timeout 600 python -m pytest tests -q -m gpu -s -k "heldout_perplexity or trajectory_at_baseline or compat_trajectory or trajectory_matches_reference or trajectory_within or full_parallelism or sharded_trainer" 2>&1 | grep -E "rel|LL/token|held-out|passed|failed" | cut -c1-400 > gpurun_out/r2_margins.log; cat gpurun_out/r2_margins.log
