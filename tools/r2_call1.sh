# round 2, first GPU call: full GPU test suite (not -x: every failure is wanted), the default bench line, launch list + ncu capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/r2_pytest_gpu_1.log 2>&1; tail -25 gpurun_out/r2_pytest_gpu_1.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_acm1m_a.json 2> gpurun_out/r2_bench_acm1m_a.err; tail -c 1500 gpurun_out/r2_bench_acm1m_a.json; tail -3 gpurun_out/r2_bench_acm1m_a.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_a.json 2>&1; tail -c 600 gpurun_out/r2_bench_ref_a.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_acm1m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
RUN_ONE_RING=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 8 -c 2 -o gpurun_out/prof_sweep_r2_acm1m -f python tools/run_one.py acm_2v 6 1000000 > gpurun_out/ncu_full_acm1m.log 2>&1
tail -3 gpurun_out/ncu_full_acm1m.log
ls -la gpurun_out/*.ncu-rep | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
