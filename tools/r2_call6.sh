set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "frozen_sweep_tracks or baseline_shapes" > gpurun_out/r2_bkt_mirror.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -6 gpurun_out/r2_bkt_mirror.log
[ $rc -eq 124 ] && { echo HANG; exit 1; }
for mode in 1 0; do
MVTM_RING=1 MVTM_BUCKETED=$mode timeout 300 python tools/run_one.py acm_2v 40 400000 > gpurun_out/r2_bkt_ab_acm_$mode.log 2>&1; echo "acm mode $mode rc=$?"; grep -E "^(1|5|10|20|30|40) " gpurun_out/r2_bkt_ab_acm_$mode.log
MVTM_RING=1 MVTM_BUCKETED=$mode timeout 300 python tools/run_one.py lda_100k 40 > gpurun_out/r2_bkt_ab_lda_$mode.log 2>&1; echo "lda mode $mode rc=$?"; grep -E "^(1|5|10|20|30|40) " gpurun_out/r2_bkt_ab_lda_$mode.log
done
