"""Sweep the launch knobs (warps per CTA, ring depth) on one workload; prints ms/sweep and tokens/s."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvtopicmodel_b200 import Engine, corpus

def run(name, cfg, grid, sweeps=6, warm=2):
    K, Vs, views = cfg if isinstance(cfg, tuple) else corpus.generate(cfg)
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
                          "Gtok_s": round(ntok / np.mean(ms) / 1e6, 3), "viol": bad, "ms_view": [round(x, 3) for x in e.stats()["ms_view"]],
                          "tok_view": e.ntok}), flush=True)
        e.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "lda"
    if which in ("lda", "all"):
        run("lda_100k", "lda_100k", [(0, 2), (0, 1), (16, 1), (12, 1)])
    if which in ("uniform",):
        run("uniform_k1000_v400k", corpus.generate_uniform(100_000, 1000, 400_000, 200), [(0, 1), (0, 2), (0, 3)], sweeps=4, warm=1)
        run("uniform_k500_v400k", corpus.generate_uniform(100_000, 500, 400_000, 200), [(0, 1), (0, 2), (0, 3)], sweeps=4, warm=1)
    if which in ("stress",):
        cfg = dict(corpus.CONFIGS["stress_4v"]); cfg["D"] = 60_000
        run("stress_4v_60k", cfg, [(0, 0)], sweeps=4, warm=4)
    if which in ("acmtext",):
        cfg = dict(D=400_000, K=1000, views=[(100_000, 120, 0.5, 1.0, 1024)])
        run("acm_text_only", cfg, [(0, 0)], sweeps=4, warm=1)
    if which in ("acm",):
        run("acm_2v", "acm_2v", [(0, 0), (16, 1), (12, 3)], sweeps=4, warm=1)
    if which in ("pubmed",):
        cfg = dict(corpus.CONFIGS["pubmed_3v"]); cfg["D"] = 250_000
        run("pubmed_3v_quarter", cfg, [(0, 0), (16, 1)], sweeps=4, warm=1)
    if which in ("k1000", "all"):
        cfg = dict(D=100_000, K=1000, views=[(200_000, 200, 0.6, 1.0, 2048)])
        run("k1000_v200k", cfg, [(0, 2), (0, 1)])
