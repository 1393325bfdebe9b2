# round 2: the HBM-bound control (uniform words, K = 1000, 1.6 GB table) on the DIRECT kernel: bench line + one --set full capture
set -x
mkdir -p gpurun_out
timeout 200 python bench.py --workload uniform_k1000 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2_bench_final_uniform_k1000.json 2> gpurun_out/r2_bench_final_uniform_k1000.err; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_final_uniform_k1000.json | cut -c1-250
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 4 -c 1 -o gpurun_out/prof_sweep_r2_final_uniform -f python tools/run_one.py uniform_k1000 5 > gpurun_out/ncu_full_final_uniform.log 2>&1; tail -2 gpurun_out/ncu_full_final_uniform.log
