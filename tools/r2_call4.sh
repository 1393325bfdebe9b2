# round 2, 8-GPU call: torch-free driver at world 2/4/8, bench at N=8 (the BASELINE target: 1 M docs, 2 views, K = 1000) and N=4
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_multi8.log 2>&1; echo "multi rc=$?"; tail -6 gpurun_out/r2_multi8.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_acm1m_${n}gpu.json 2> gpurun_out/r2_bench_acm1m_${n}gpu.err; echo "bench$n rc=$?"; tail -n 1 gpurun_out/r2_bench_acm1m_${n}gpu.json | cut -c1-330; tail -n 1 gpurun_out/r2_bench_acm1m_${n}gpu.json | grep -o '"e2e": {[^}]*}' | cut -c1-300
done
