# round 2, final 8-GPU evidence: torch-free driver, the default bench line at N = 8, BASELINE configs[3] and configs[4] at N = 8
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_multi_final.log 2>&1; echo "multi rc=$?"; tail -4 gpurun_out/r2_multi_final.log
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 20 --warmup 5 $2 > gpurun_out/r2_bench_final_8gpu_$1.json 2> gpurun_out/r2_bench_final_8gpu_$1.err; echo "$1 rc=$?"
  tail -n 1 gpurun_out/r2_bench_final_8gpu_$1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3),'G tok/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value']/1e9,3), 'whole-job frac', round(d['roofline']['whole_job_frac'],3), d['config'].get('invariant_violations'))"
}
run acm1m "" 29601
run acm1m_reserve4 "--reserve-sms 4" 29602
run pubmed1m "--workload pubmed_3v" 29603
run stress2m "--workload stress_4v --steps 10 --warmup 3" 29604
