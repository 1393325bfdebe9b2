"""A log-likelihood TRAJECTORY of the reference implementation itself: the jar's sampler + updater + modelLogLikelihood
(see make_reference_sampler_vectors.py for what is executed and what is shimmed) run for 30 sweeps over a 400-document two-view
corpus with the burn-in ramp of p_a (M:1166-1169).  Output: tests/golden/reference_trajectory.json -- corpus, initial assignments
and LL per view every 5 sweeps.  north_star check (c) compares the engine's trajectory with it (1 %).  Takes a few minutes
(a Python bytecode interpreter runs ~360 K token draws).  Needs /root/reference."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_reference_sampler_vectors as G  # noqa: E402
from mvtopicmodel_b200 import corpus  # noqa: E402
from oracle import oracle as O  # noqa: E402

CFG = dict(D=400, K=20, views=[(300, 25, 0.5, 1.0, 128), (60, 5, 0.5, 0.8, 32)])
SEED, SWEEPS, EVERY = 77, 30, 5


def main():
    K, Vs, views = corpus.generate(CFG)
    M = len(Vs)
    o = O.Oracle(K, Vs, views, seed=SEED)
    o.init_assignments()
    z0 = [o.get_assignments(m).tolist() for m in range(M)]
    alpha = np.full((M, K + 1), 0.1)
    ref = G.RefSampler(K, Vs, views, z0, SEED, alpha, np.full(M, 0.1 * K), np.full(M, 0.01), np.ones(M), np.full((M, M), 0.2), np.ones((M, M)))
    out = {"cfg": CFG, "seed": SEED, "K": K, "V": Vs, "views": [{"off": v[0].tolist(), "word": v[1].tolist()} for v in views], "z0": z0,
           "loglik": [[0, G.reference_loglik(ref)]]}
    t0 = time.time()
    for it in range(1, SWEEPS + 1):
        pa = min(it / 100.0 + 0.3, 1.1)                                   # M:1166-1169
        for row in ref.worker.fields["p_a"]:
            for j in range(M):
                row[j] = pa
        ref.sweep(it)
        if it % EVERY == 0:
            out["loglik"].append([it, G.reference_loglik(ref)])
            print(it, out["loglik"][-1][1], "%.0f s" % (time.time() - t0), flush=True)
    out["z_final"] = [list(z) for z in ref.z]
    out["counters"] = dict(ref.counters)
    json.dump(out, open(os.path.join(HERE, "reference_trajectory.json"), "w"))
    print("tokens", [len(z) for z in z0], ref.counters, "bytecode steps", ref.vm.steps)


if __name__ == "__main__":
    main()
