# round 2: DIRECT kernel at K = 512 with 128 registers (two documents per warp, 16 warps per SM) -- parity subset under MVTM_DIRECT=1, A/B vs the ring
set -x
mkdir -p gpurun_out
cp mvtopicmodel_b200/libmvtm.so /tmp/keep.so; cp build_ab/libmvtm_f1.so mvtopicmodel_b200/libmvtm.so
MVTM_DIRECT=1 timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "direct_kernel or frozen_sweep or count_invariants or conditionals_on_frozen or full_size_properties_lda or baseline_shapes" > gpurun_out/r2_direct_k512_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_direct_k512_pytest.log
cp /tmp/keep.so mvtopicmodel_b200/libmvtm.so
AB_REPS=1 timeout 400 python tools/ab.py mvtopicmodel_b200/libmvtm.so build_ab/libmvtm_f1.so@MVTM_DIRECT=1 lda_100k > gpurun_out/r2_ab_direct_k512.log 2>&1; cat gpurun_out/r2_ab_direct_k512.log
