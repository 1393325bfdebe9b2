# round 2, third GPU call (2 GPUs): NCCL behind the C ABI -- torch-free C++ driver, bench at N=2 through mvtm_sweep_dist /
# mvtm_sweep_host_dist -- and the new BASELINE-shape parity tests
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_multi.log 2>&1; echo "multi rc=$?"; tail -15 gpurun_out/r2_multi.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "baseline_shapes or ring_ or tolerate_unassigned" --durations=8 > gpurun_out/r2_shapes.log 2>&1; echo "shapes rc=$?"; tail -25 gpurun_out/r2_shapes.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_acm1m_2gpu.json 2> gpurun_out/r2_bench_acm1m_2gpu.err; echo "bench2 rc=$?"; tail -c 1800 gpurun_out/r2_bench_acm1m_2gpu.json; tail -5 gpurun_out/r2_bench_acm1m_2gpu.err
