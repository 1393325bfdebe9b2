"""How much of a long-tail pass is the tail?  stress_4v's text view (log-normal lengths up to 16 K tokens, K = 2000) against the same
view with the lengths clipped at 2 K: if the pass time per token is the same, the longest-first work list hides the outliers and a
CTA-per-document path would buy nothing.  usage: python tools/tail_exp.py [docs]"""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mvtopicmodel_b200 import Engine, corpus
docs = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
for nmax in (16384, 2048):
    cfg = copy.deepcopy(corpus.CONFIGS["stress_4v"])
    v = cfg["views"][0]
    cfg["views"][0] = (v[0], v[1], v[2], v[3], nmax)
    K, Vs, views = corpus.generate(cfg, docs=docs)
    lens = views[0][0][1:] - views[0][0][:-1]
    e = Engine(K, Vs, views, seed=1, ring_depth=1); e.init_assignments()
    ms = []
    for it in range(1, 9):
        e.sweep(it)
        if it > 3:
            ms.append(e.stats()["ms_view"])
    ms = np.array(ms).mean(0)
    print(f"Nmax {nmax}: longest doc {int(lens.max())}, tokens/view {e.ntok}, ms/view {ms.round(2).tolist()}, "
          f"text pass {e.ntok[0] / ms[0] / 1e6:.3f} G tok/s, docs > 2048 tokens: {int((lens > 2048).sum())} holding {int(lens[lens > 2048].sum())} tokens", flush=True)
    e.close()
