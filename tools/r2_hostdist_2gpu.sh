# round 2: mvtm_sweep_host_dist with the compare-and-keep path -- multi-GPU tests (torch-free driver at world 2, single-rank comm test) and the
# bench at 2 GPUs (e2e leg = mvtm_sweep_host_dist)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2_multi_hostdist.log 2>&1; echo "multi rc=$?"; tail -6 gpurun_out/r2_multi_hostdist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_hostdist_2gpu.json 2> gpurun_out/r2_bench_hostdist_2gpu.err; echo "bench rc=$?"
tail -n 1 gpurun_out/r2_bench_hostdist_2gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,3),'G tok/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value']/1e9,3), round(d['e2e']['ms_per_step'],2), 'ms', d['config'].get('invariant_violations'))"
MVTM_HOST_COMPARE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_hostdist_2gpu_nocompare.json 2> gpurun_out/r2_bench_hostdist_2gpu_nocompare.err; echo "bench rc=$?"
tail -n 1 gpurun_out/r2_bench_hostdist_2gpu_nocompare.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('recount path:', round(d['value']/1e9,3),'G tok/s', round(d['ms_per_step'],2),'ms; e2e', round(d['e2e']['value']/1e9,3), round(d['e2e']['ms_per_step'],2), 'ms', d['config'].get('invariant_violations'))"
tail -3 gpurun_out/r2_bench_hostdist_2gpu.err
