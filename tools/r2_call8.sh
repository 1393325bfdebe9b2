set -x
mkdir -p gpurun_out
MVTM_RING=1 timeout 900 python tools/ab.py build_ab/libmvtm_base.so build_ab/libmvtm_v0.so build_ab/libmvtm_v1.so acm_2v:400000 pubmed_3v:125000 > gpurun_out/r2_ab_variants.log 2>&1; cat gpurun_out/r2_ab_variants.log
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/r2_pytest_gpu_3.log 2>&1; echo "suite rc=$?"; tail -12 gpurun_out/r2_pytest_gpu_3.log
