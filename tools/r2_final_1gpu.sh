# round 2, final 1-GPU evidence: the whole GPU suite, the default bench line + reference arm, the ncu launch list of the bench
# command and one --set full capture of the dominant kernel at the bench size
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "suite rc=$?"; tail -14 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2>&1; tail -n 1 gpurun_out/r2_bench_ref_final.json | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final_1gpu.err; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_final_1gpu.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_final.log 2>&1
RUN_ONE_RING=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 8 -c 2 -o gpurun_out/prof_sweep_r2_final_acm1m -f python tools/run_one.py acm_2v 6 1000000 > gpurun_out/ncu_full_final.log 2>&1; tail -2 gpurun_out/ncu_full_final.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
