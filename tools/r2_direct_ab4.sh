# round 2: DIRECT kernel v2 (om / cpar / gaf addressed from the base registers, no __syncwarp in the token loop) vs v1
set -x
mkdir -p gpurun_out
AB_REPS=1 timeout 1200 python tools/ab.py mvtopicmodel_b200/libmvtm.so build_ab/libmvtm_e1.so acm_2v:200000 pubmed_3v:60000 stress_4v:100000 acmtext > gpurun_out/r2_ab_direct4.log 2>&1
cat gpurun_out/r2_ab_direct4.log
