# round 2: first run of the bucketed sampler -- targeted tests under short timeouts (a new hot kernel can hang), then A/B on the bench
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frozen_sweep_tracks" > gpurun_out/r2_bkt_mirror.log 2>&1; rc=$?; echo "mirror rc=$rc"; tail -12 gpurun_out/r2_bkt_mirror.log
[ $rc -eq 124 ] && { echo HANG; exit 1; }
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "invariants or baseline_shapes or inference_matches" > gpurun_out/r2_bkt_shapes.log 2>&1; rc=$?; echo "shapes rc=$rc"; tail -12 gpurun_out/r2_bkt_shapes.log
[ $rc -eq 124 ] && { echo HANG; exit 1; }
for mode in 0 1; do
MVTM_BUCKETED=$mode timeout 300 python tools/run_one.py acm_2v 12 400000 > gpurun_out/r2_bkt_ab_acm_$mode.log 2>&1; echo "acm mode $mode rc=$?"; tail -4 gpurun_out/r2_bkt_ab_acm_$mode.log
MVTM_BUCKETED=$mode timeout 300 python tools/run_one.py lda_100k 12 > gpurun_out/r2_bkt_ab_lda_$mode.log 2>&1; echo "lda mode $mode rc=$?"; tail -3 gpurun_out/r2_bkt_ab_lda_$mode.log
done
