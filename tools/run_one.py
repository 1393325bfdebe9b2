"""Run a few sweeps of one workload (for ncu captures): python tools/run_one.py WORKLOAD [sweeps] [total_docs]
WORKLOAD: any bench.py workload (acm_2v, lda_100k, pubmed_3v, stress_4v, uniform_k1000) or the short names lda / acm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
which = {"lda": "lda_100k", "acm": "acm_2v"}.get(sys.argv[1], sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
docs = int(sys.argv[3]) if len(sys.argv) > 3 else bench.DEFAULT_DOCS.get(which, 100_000)
K, Vs, views = bench.build_corpus(which, docs, 0, 1)
from mvtopicmodel_b200 import Engine
e = Engine(K, Vs, views, seed=1, ring_depth=int(os.environ.get("RUN_ONE_RING", "0"))); e.init_assignments()
print("tokens per view", e.ntok, flush=True)
for it in range(1, n + 1):
    e.sweep(it)
    st = e.stats()
    print(it, st["ms_total"], st["ms_view"], st["ring_depth"], flush=True)
assert e.check_invariants() == 0
