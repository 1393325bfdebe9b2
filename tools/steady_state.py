"""Throughput over a long run: the bench times sweeps 12-31 after a random initialisation (96 % of the tokens still change their
topic); this prints pass times and the changed fraction every 25 sweeps up to `sweeps`, on a slice of a BASELINE shape.
usage: python tools/steady_state.py [workload] [docs] [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mvtopicmodel_b200 import Engine
wl = sys.argv[1] if len(sys.argv) > 1 else "acm_2v"
docs = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
K, Vs, views = bench.build_corpus(wl, docs, 0, 1)
e = Engine(K, Vs, views, seed=1, ring_depth=1); e.init_assignments()
ntok = sum(e.ntok)
print(f"{wl}: {docs} docs, tokens {e.ntok}, K={K}", flush=True)
for it in range(1, sweeps + 1):
    e.sweep(it)
    if it in (1, 5, 10, 20, 30) or it % 25 == 0:
        st = e.stats()
        print(f"sweep {it:4d}  ms/view {[round(x, 3) for x in st['ms_view']]}  {ntok / sum(st['ms_view']) / 1e6:.3f} G tok/s  changed {st['changed'] / st['tokens']:.3f}  "
              f"LL/token {(e.loglik() / [max(1, n) for n in e.ntok]).round(4).tolist()}", flush=True)
assert e.check_invariants() == 0
