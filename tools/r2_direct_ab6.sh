# round 2: DIRECT kernel with packed fp32 pairs (fma.f32x2) and beta folded into the chunk sum (no per-chunk beta*sum(q) registers) vs v2
set -x
mkdir -p gpurun_out
AB_REPS=1 timeout 1200 python tools/ab.py mvtopicmodel_b200/libmvtm.so build_ab/libmvtm_e3.so acm_2v:200000 pubmed_3v:60000 stress_4v:100000 uniform_k1000 > gpurun_out/r2_ab_direct6.log 2>&1
cat gpurun_out/r2_ab_direct6.log
