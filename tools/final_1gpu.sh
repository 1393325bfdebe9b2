set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/final_lda.log 2>&1; tail -1 gpurun_out/final_lda.log | cut -c1-200
python bench.py --workload acm_2v --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final_acm.log 2>&1; tail -1 gpurun_out/final_acm.log | cut -c1-200
python bench.py --workload pubmed_3v --docs 125000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final_pubmed125k.log 2>&1; tail -1 gpurun_out/final_pubmed125k.log | cut -c1-200
python bench.py --workload stress_4v --docs 100000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_stress100k.log 2>&1; tail -1 gpurun_out/final_stress100k.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.log 2>&1; tail -1 gpurun_out/final_ref.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_v9.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 5 -c 1 -o gpurun_out/prof_sweep_r1_v9_lda -f python tools/run_one.py lda 6 > gpurun_out/ncu_full_lda.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_view -s 10 -c 2 -o gpurun_out/prof_sweep_r1_v9_acm -f python tools/run_one.py acm 6 > gpurun_out/ncu_full_acm.log 2>&1
python tools/sanitize_case.py 2>&1 | tail -1
python __graft_entry__.py --smoke 2>&1 | tail -1
