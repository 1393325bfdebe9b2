# round 2: host step verdict on the wide communicator from the handle's stream (not behind the CTA-limited exchanges): multi-GPU tests + a
# three-view bench at 2 GPUs (e2e leg = mvtm_sweep_host_dist), with and without the comparison
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2_multi_hostdist2.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/r2_multi_hostdist2.log
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --workload pubmed_3v --docs 300000 > gpurun_out/r2_bench_hostdist2_$1.json 2> gpurun_out/r2_bench_hostdist2_$1.err; echo "bench rc=$?"
  tail -n 1 gpurun_out/r2_bench_hostdist2_$1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e9,3),'G tok/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value']/1e9,3), round(d['e2e']['ms_per_step'],2), 'ms', d['config'].get('invariant_violations'))"
}
run compare 29621
MVTM_HOST_COMPARE=0 run recount 29622
