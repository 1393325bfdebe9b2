"""Sweep the launch knobs (warps per CTA, ring depth) on one workload; prints ms/sweep and tokens/s."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvtopicmodel_b200 import Engine, corpus

def run(name, cfg, grid, sweeps=6, warm=2):
    K, Vs, views = corpus.generate(cfg)
    ntok = sum(len(v[1]) for v in views)
    print(f"== {name}: K={K} V={Vs} tokens={ntok}", flush=True)
    for (W, R) in grid:
        e = Engine(K, Vs, views, seed=1, warps_per_cta=W, ring_depth=R)
        e.init_assignments()
        ms = []
        for it in range(1, warm + sweeps + 1):
            e.sweep(it)
            if it > warm:
                ms.append(e.stats()["ms_total"])
        bad = e.check_invariants()
        print(json.dumps({"cfg": name, "W": W, "R": R, "ms": round(float(np.mean(ms)), 3), "min_ms": round(float(np.min(ms)), 3),
                          "Gtok_s": round(ntok / np.mean(ms) / 1e6, 3), "viol": bad}), flush=True)
        e.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "lda"
    if which in ("lda", "all"):
        run("lda_100k", "lda_100k", [(16, 4), (19, 4), (20, 3), (23, 3), (24, 2), (16, 2), (12, 4), (8, 4)])
    if which in ("k1000", "all"):
        cfg = dict(D=100_000, K=1000, views=[(200_000, 200, 0.6, 1.0, 2048)])
        run("k1000_v200k", cfg, [(16, 2), (12, 3), (10, 4), (8, 5), (8, 3), (6, 6), (16, 1)])
