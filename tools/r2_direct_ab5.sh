# round 2: DIRECT kernel, chunk picked by a branch on the selected lane's chunk index (switch) instead of the select tree; new kernel-equality test
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "direct_kernel or ring_depths or ring_autotune or frozen_sweep" > gpurun_out/r2_direct_pytest5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_direct_pytest5.log
AB_REPS=1 timeout 1200 python tools/ab.py mvtopicmodel_b200/libmvtm.so build_ab/libmvtm_e2.so acm_2v:200000 pubmed_3v:60000 stress_4v:100000 > gpurun_out/r2_ab_direct5.log 2>&1
cat gpurun_out/r2_ab_direct5.log
