set -x
mkdir -p gpurun_out
MVTM_RING=1 timeout 900 python tools/ab.py build_ab/libmvtm_base.so build_ab/libmvtm_v7.so build_ab/libmvtm_v9.so acm_2v:200000 pubmed_3v:60000 lda_100k stress_4v:40000 > gpurun_out/r2_ab_variants4.log 2>&1; cat gpurun_out/r2_ab_variants4.log
