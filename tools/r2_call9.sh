set -x
mkdir -p gpurun_out
MVTM_RING=1 timeout 900 python tools/ab.py build_ab/libmvtm_base.so build_ab/libmvtm_v1.so build_ab/libmvtm_v2.so build_ab/libmvtm_v4.so acm_2v:200000 pubmed_3v:60000 > gpurun_out/r2_ab_variants2.log 2>&1; cat gpurun_out/r2_ab_variants2.log
timeout 200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k single_rank 2>&1 | tail -5
