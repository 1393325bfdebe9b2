# round 2, second GPU call: the reference-compat mode (Q1 index, MALLET Beta) and the inference pin, each test under its own
# short timeout (the first attempt at the Q1 mode hung a box in round 1), then the whole GPU suite.
set -x
mkdir -p gpurun_out
for t in test_q1_frozen_sweep_tracks_oracle_mirror test_q1_live_sweeps_keep_invariants_and_limits test_compat_trajectory_matches_reference_bytecode test_mallet_beta_flag_changes_coupling_only_above_one; do
    timeout 150 python -m pytest tests/test_gpu_compat.py -x -q -k "$t" > gpurun_out/r2_compat_$t.log 2>&1
    rc=$?
    echo "== $t rc=$rc"; tail -4 gpurun_out/r2_compat_$t.log
    [ $rc -eq 124 ] && { echo "HANG: stopping"; exit 1; }
done
timeout 300 python -m pytest tests/test_gpu_bytecode.py -x -q -m gpu > gpurun_out/r2_bytecode.log 2>&1; echo "bytecode rc=$?"; tail -5 gpurun_out/r2_bytecode.log
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2_pytest_gpu_2.log 2>&1; echo "suite rc=$?"; tail -16 gpurun_out/r2_pytest_gpu_2.log
