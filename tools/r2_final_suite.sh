# round 2: the whole GPU suite + smoke on the final binary
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu_final3.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_final3.log
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
