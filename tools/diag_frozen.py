"""Diagnostic: where do the frozen engine sweep and the fp64 mirror disagree, and how close was the uniform to a
boundary there?  (fp32 rounding => margins ~1e-7 of the total mass.)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import random_corpus
from mvtopicmodel_b200 import Engine
from oracle import oracle as O

K, Vs, means = 500, [800], [40]
views = random_corpus(K + 3, 400, K, Vs, means)
seed = 77
e = Engine(K, Vs, views, seed=seed); o = O.Oracle(K, Vs, views, seed=seed)
e.init_assignments(); o.init_assignments()
z0 = o.get_assignments(0).copy()
nwk, nk = o.get_counts(0)
e.sweep(1, update_global=False); o.sweep(1, O.F_ENGINE_MIRROR | O.F_FROZEN)
ze, zo = e.get_assignments(0), o.get_assignments(0)
off, w = views[0]
J = (K + 127) // 128
order = np.array([4 * ((i >> 2) // J + 32 * ((i >> 2) % J)) + (i & 3) for i in range(J * 128)])
order = order[order < K]
roots = 0; tot_bad = int((ze != zo).sum())
print("tokens", len(ze), "disagree", tot_bad)
for d in range(len(off) - 1):
    b, en = off[d], off[d + 1]
    bad = np.nonzero(ze[b:en] != zo[b:en])[0]
    if len(bad) == 0:
        continue
    roots += 1
    pos = int(bad[0])
    # state before token pos in the oracle's chain == engine's chain (first disagreement)
    zz = z0[b:en].copy(); zz[:pos] = zo[b:b + pos]
    nd = np.bincount(zz, minlength=K).astype(np.float64); nd[z0[b + pos]] -= 1
    wt = (nwk[w[b + pos]] + 0.01) * (nd + 0.1) / (nk + 0.01 * Vs[0])
    x = O.philox([pos, d, 1, 0], [seed & 0xffffffff, seed >> 32])
    u = (int(x[0]) >> 8) / 2.0**24
    cum = np.cumsum(wt[order]); s = u * cum[-1]
    idx = int(np.searchsorted(cum, s, side="right"))
    margin = min(abs(cum[idx] - s), abs(s - (cum[idx - 1] if idx else 0))) / cum[-1]
    print(f"doc {d} len {en-b} pos {pos} followups {len(bad)-1} oracle {zo[b+pos]} engine {ze[b+pos]} numpy {order[idx]} margin {margin:.3e} u {u:.7f} "
          f"order-dist {abs(int(np.where(order==ze[b+pos])[0][0]) - int(np.where(order==zo[b+pos])[0][0]))}")
print("root docs", roots)
