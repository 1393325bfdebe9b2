# round 2: DIRECT sweep kernel, register cap / warps per SM (168 -> 12 warps ... 136 -> 15 warps), single-view and HBM-bound workloads
set -x
mkdir -p gpurun_out
AB_REPS=1 timeout 1200 python tools/ab.py build_ab/libmvtm_vD3.so build_ab/libmvtm_r168.so@MVTM_DIRECT=1 build_ab/libmvtm_r152.so@MVTM_DIRECT=1 build_ab/libmvtm_r144.so@MVTM_DIRECT=1 build_ab/libmvtm_r136.so@MVTM_DIRECT=1 acm_2v:200000 acmtext uniform_k1000 > gpurun_out/r2_ab_direct2.log 2>&1
cat gpurun_out/r2_ab_direct2.log
