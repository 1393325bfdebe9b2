# round 2: first GPU run of the DIRECT sweep kernel (n_wk rows in registers): parity subset with MVTM_DIRECT=1, then A/B vs the committed build
set -x
mkdir -p gpurun_out
MVTM_DIRECT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "frozen_sweep or count_invariants or full_size or baseline_shapes or sweep_host or edge_cases or ring_depths" > gpurun_out/r2_direct_pytest1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_direct_pytest1.log
AB_REPS=1 timeout 900 python tools/ab.py build_ab/libmvtm_vD3.so build_ab/libmvtm_d64.so@MVTM_DIRECT=1 build_ab/libmvtm_d104.so@MVTM_DIRECT=1 acm_2v:200000 lda_100k stress_4v:40000 pubmed_3v:60000 > gpurun_out/r2_ab_direct1.log 2>&1
cat gpurun_out/r2_ab_direct1.log
