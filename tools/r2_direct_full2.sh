# round 2: packed chunk weights (fma.f32x2, no per-chunk beta*sum(q) registers) in EVERY sweep kernel and the probe -- whole GPU suite, then A/B
# vs the round-start build (vD3) on the ring-kernel shapes and vs the previous DIRECT build on the K = 1000 shape
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/r2_pytest_gpu_packed.log 2>&1; echo "suite rc=$?"; tail -10 gpurun_out/r2_pytest_gpu_packed.log
AB_REPS=1 timeout 1200 python tools/ab.py build_ab/libmvtm_vD3.so mvtopicmodel_b200/libmvtm.so lda_100k small_3v acm_2v:200000 > gpurun_out/r2_ab_packed.log 2>&1
cat gpurun_out/r2_ab_packed.log
