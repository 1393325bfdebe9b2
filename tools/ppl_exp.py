import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from mvtopicmodel_b200 import Engine, corpus
from mvtopicmodel_b200.model import split_for_completion
from oracle import oracle as O
O.build()
K, Vs, views = corpus.generate("small_3v")
D = len(views[0][0]) - 1
cut = 2400
def part(lo, hi):
    return [(np.ascontiguousarray(off[lo:hi + 1] - off[lo]), np.ascontiguousarray(w[off[lo]:off[hi]])) for off, w in views]
train, held = part(0, cut), part(cut, D)
obs, ev = split_for_completion(held)
def ppl_engine(counts, seed):
    f = Engine(K, Vs, obs, seed=seed)
    for m in range(3): f.set_counts(m, *counts[m])
    f.init_assignments_from_counts()
    for it in range(1, 11): f.sweep(it, update_global=0)
    return np.array([np.exp(-(lambda r: r[0] / r[1])(f.heldout_loglik(m, ev[m][0], ev[m][1]))) for m in range(3)])
def ppl_oracle(counts, seed):
    f = O.Oracle(K, Vs, obs, seed=seed)
    for m in range(3): f.set_counts(m, *counts[m])
    f.init_from_phi()
    for it in range(1, 11): f.sweep(it, O.F_FROZEN)
    out = []
    for m in range(3):
        ll, n = O.heldout_loglik(obs[m], f.get_assignments(m), counts[m][0], counts[m][1], ev[m], np.full(K, 0.1), 0.01, 0.01 * Vs[m])
        out.append(np.exp(-ll / n))
    return np.array(out)
for flags in (0, O.F_STALE_TREES):
    e = Engine(K, Vs, train, seed=21, max_ctas=16, warps_per_cta=4); e.init_assignments()
    o = O.Oracle(K, Vs, train, seed=21); o.init_assignments(); o.rebuild_trees()
    it = 0
    for stop in (20, 40, 60, 100, 140, 200):
        while it < stop:
            it += 1; e.sweep(it); o.sweep(it, flags)
        ce = [e.get_counts(m) for m in range(3)]; co = [o.get_counts(m) for m in range(3)]
        print(flags, stop, "E/eng", ppl_engine(ce, 5).round(1), "E/eng s2", ppl_engine(ce, 6).round(1), "E/orc", ppl_oracle(ce, 5).round(1),
              "O/orc", ppl_oracle(co, 5).round(1), "O/eng", ppl_engine(co, 5).round(1), flush=True)
