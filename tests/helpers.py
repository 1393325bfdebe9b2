import numpy as np


def random_corpus(seed, D, K, Vs, mean_lens, empty_frac=0.1, oov=False):
    """Ragged doc-aligned CSR views with empty documents, single-token documents and (optionally) OOV ids."""
    rng = np.random.default_rng(seed)
    views = []
    for V, ml in zip(Vs, mean_lens):
        lens = rng.poisson(ml, size=D).astype(np.int64)
        lens[rng.random(D) < empty_frac] = 0
        lens[rng.random(D) < 0.05] = 1
        off = np.zeros(D + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        # zipf-ish words so that hot rows collide across warps
        w = np.minimum((rng.pareto(1.1, size=int(off[-1])) * 3).astype(np.int64), V - 1).astype(np.int32)
        if oov and len(w) > 10:
            w[rng.integers(0, len(w), size=max(1, len(w) // 50))] = V + 3
        views.append((off, w))
    return views


def recount(views, zs, K, Vs):
    out = []
    for (off, w), z, V in zip(views, zs, Vs):
        nwk = np.zeros((V, K), dtype=np.int64)
        ok = (w >= 0) & (w < V) & (z >= 0)
        np.add.at(nwk, (w[ok], z[ok]), 1)
        nk = np.bincount(z[z >= 0], minlength=K)
        out.append((nwk, nk))
    return out
