# round 2: DIRECT sweep kernel as the K = 1024 default -- whole GPU suite, default bench line, ncu launch list + one --set full capture
# of the text pass at the bench size, K = 2048 on/off on a corpus large enough that the long-tail documents do not bound the pass
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2_pytest_gpu_direct.log 2>&1; echo "suite rc=$?"; tail -14 gpurun_out/r2_pytest_gpu_direct.log
timeout 900 python bench.py > gpurun_out/r2_bench_direct_1gpu.json 2> gpurun_out/r2_bench_direct_1gpu.err; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_direct_1gpu.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_direct.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_direct.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 8 -c 2 -o gpurun_out/prof_sweep_r2_direct_acm1m -f python tools/run_one.py acm_2v 6 1000000 > gpurun_out/ncu_full_direct.log 2>&1; tail -2 gpurun_out/ncu_full_direct.log
AB_REPS=1 timeout 900 python tools/ab.py mvtopicmodel_b200/libmvtm.so@MVTM_DIRECT=0 mvtopicmodel_b200/libmvtm.so@MVTM_DIRECT=1 stress_4v:200000 > gpurun_out/r2_ab_direct3_stress.log 2>&1; cat gpurun_out/r2_ab_direct3_stress.log
