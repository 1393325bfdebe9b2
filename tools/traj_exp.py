import sys, os, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvtopicmodel_b200 import Engine
from oracle import oracle as O
O.build()
g = json.load(open("tests/golden/reference_trajectory.json"))
K, Vs = g["K"], g["V"]
views = [(np.array(v["off"], dtype=np.int64), np.array(v["word"], dtype=np.int32)) for v in g["views"]]
M = len(Vs)
present = [np.ones(len(views[0][0]) - 1, dtype=np.uint8)] + [((v[0][1:] - v[0][:-1]) > 0).astype(np.uint8) for v in views[1:]]
marks = {it: np.array(ll) for it, ll in g["loglik"]}
print("reference      ", {it: np.round(ll, 0).tolist() for it, ll in marks.items()})
for seed in (77, 1, 2, 3):
    e = Engine(K, Vs, views, seed=seed, present=present, max_ctas=2, warps_per_cta=2)
    for m in range(M): e.set_assignments(m, np.array(g["z0"][m], dtype=np.int32))
    out = {}
    for it in range(1, 31):
        e.set_hyper(p_a=np.full((M, M), min(it / 100.0 + 0.3, 1.1))); e.sweep(it)
        if it in marks: out[it] = np.round(100 * (e.loglik(True) - marks[it]) / np.abs(marks[it]), 2).tolist()
    print("engine seed", seed, "rel % vs reference", out)
# the oracle with other seeds (same algorithm as the reference, different randomness): the reference's own run-to-run spread
for seed in (1, 2, 3):
    o = O.Oracle(K, Vs, views, seed=seed, present=present)
    o.set_assignments([np.array(z, dtype=np.int32) for z in g["z0"]]); o.rebuild_trees()
    out = {}
    for it in range(1, 31):
        o.set_hyper(p_a=np.full((M, M), min(it / 100.0 + 0.3, 1.1))); o.sweep(it, O.F_STALE_TREES | O.F_Q1_COMPAT)
        if it in marks: out[it] = np.round(100 * (o.loglik(True) - marks[it]) / np.abs(marks[it]), 2).tolist()
    print("oracle (reference-faithful) seed", seed, "rel %", out)
