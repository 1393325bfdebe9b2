# round 2, final 8-GPU evidence with the DIRECT kernel and the compare-and-keep host step: torch-free driver at world 8, the default bench
# line at N = 8 (two SM reservations), BASELINE configs[3] and configs[4] at N = 8
set -x
mkdir -p gpurun_out
g++ -std=c++17 -O1 -Iinclude tests/cpp/dist_driver.cpp -o /tmp/dist_driver -Lmvtopicmodel_b200 -lmvtm -lpthread -Wl,-rpath,$PWD/mvtopicmodel_b200 && timeout 300 /tmp/dist_driver 8 8 2 > gpurun_out/r2_dist_driver_world8_final.log 2>&1; echo "driver rc=$?"; tail -3 gpurun_out/r2_dist_driver_world8_final.log
run() { # name, extra args, port
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 8 --steps 20 --warmup 5 $2 > gpurun_out/r2_bench_final2_8gpu_$1.json 2> gpurun_out/r2_bench_final2_8gpu_$1.err; echo "$1 rc=$?"
  tail -n 1 gpurun_out/r2_bench_final2_8gpu_$1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3),'G tok/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value']/1e9,3), round(d['e2e']['ms_per_step'],2), 'ms; whole-job frac', round(d['roofline']['whole_job_frac'],3), d['config'].get('invariant_violations'))"
}
run acm1m "" 29601
run acm1m_reserve4 "--reserve-sms 4" 29602
run pubmed1m "--workload pubmed_3v" 29603
run stress2m "--workload stress_4v --steps 10 --warmup 3" 29604
