"""Run a few sweeps of one workload (for ncu captures): python tools/run_one.py lda|k1000 [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvtopicmodel_b200 import Engine, corpus
which = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = {"lda": "lda_100k", "acm": "acm_2v", "k1000": dict(D=100_000, K=1000, views=[(200_000, 200, 0.6, 1.0, 2048)])}[which]
K, Vs, views = corpus.generate(cfg)
e = Engine(K, Vs, views, seed=1); e.init_assignments()
for it in range(1, n + 1):
    e.sweep(it)
    print(it, e.stats()["ms_total"], flush=True)
assert e.check_invariants() == 0
