# round 2, final 1-GPU evidence (DIRECT kernel + packed chunk weights + compare-and-keep host step): the whole GPU suite, the reference arm and the
# default bench line, the ncu launch list of the bench command and one --set full capture of the dominant kernel at the bench size, smoke
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r2_pytest_gpu_final2.log 2>&1; echo "suite rc=$?"; tail -12 gpurun_out/r2_pytest_gpu_final2.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final2.json 2>&1; tail -n 1 gpurun_out/r2_bench_ref_final2.json | cut -c1-160
timeout 600 python bench.py > gpurun_out/r2_bench_final2_1gpu.json 2> gpurun_out/r2_bench_final2_1gpu.err; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_final2_1gpu.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_final2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 8 -c 1 -o gpurun_out/prof_sweep_r2_final2_acm1m -f python tools/run_one.py acm_2v 5 1000000 > gpurun_out/ncu_full_final2.log 2>&1; tail -2 gpurun_out/ncu_full_final2.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
