set -x
for g in 16 32; do MVTM_GROUP=$g python tools/ab.py mvtopicmodel_b200/libmvtm.so acm_2v:100000 2>&1 | sed "s/^/G=$g /"; done
for w in stress_4v:40000; do python tools/ab.py mvtopicmodel_b200/libmvtm.so $w 2>&1 | head -1; done
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/sanitize_memcheck2.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck2.log
