set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frozen_sweep_tracks or bucketed_sampler" > gpurun_out/r2_bkt_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -6 gpurun_out/r2_bkt_tests.log
[ $rc -eq 124 ] && { echo HANG; exit 1; }
MVTM_RING=1 timeout 900 python tools/ab.py build_ab/libmvtm_base.so mvtopicmodel_b200/libmvtm.so acm_2v:400000 lda_100k pubmed_3v:125000 > gpurun_out/r2_ab_base_vs_now.log 2>&1; cat gpurun_out/r2_ab_base_vs_now.log
