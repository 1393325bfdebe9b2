# ncu evidence for the sweep kernel (run on one B200 under gpurun): launch list of the bench command + one --set full capture
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_15.log 2>&1; tail -3 gpurun_out/pytest_gpu_15.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
python tools/run_one.py lda 6 > gpurun_out/run_one_lda.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 5 -c 1 -o gpurun_out/prof_sweep_r1_v8_lda -f python tools/run_one.py lda 6 > gpurun_out/ncu_full_lda.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 10 -c 1 -o gpurun_out/prof_sweep_r1_v8_acm_text -f python tools/run_one.py acm 6 > gpurun_out/ncu_full_acm.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
