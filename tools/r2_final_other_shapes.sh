# round 2: final 1-GPU bench lines of the other BASELINE shapes (configs[1]; the per-GPU share of configs[3])
mkdir -p gpurun_out
timeout 100 python bench.py --workload lda_100k --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2_bench_final_lda100k.json 2>/dev/null; tail -n 1 gpurun_out/r2_bench_final_lda100k.json | cut -c1-200
timeout 100 python bench.py --workload pubmed_3v --docs 125000 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r2_bench_final_pubmed125k.json 2>/dev/null; tail -n 1 gpurun_out/r2_bench_final_pubmed125k.json | cut -c1-200
