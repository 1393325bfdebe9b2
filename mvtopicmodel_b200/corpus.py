"""Synthetic corpora of the BASELINE.json shapes as doc-aligned CSR views.

The generator is an LDA-style process with a Zipf(1.07) base measure over words, sparse per-document
topic mixtures shared between the views of a document (80 % shared / 20 % view-specific), log-normal
document lengths clipped to [1, Nmax] and per-view coverage c_m.  It is deterministic in (config, seed, shard).
"""
import math

import numpy as np

# (V_m, mean length, sigma, coverage, Nmax) per view
CONFIGS = {
    # configs[1]: synthetic single-view LDA shape -- the 1-GPU bench workload
    "lda_100k": dict(D=100_000, K=500, views=[(50_000, 200, 0.6, 1.0, 2048)]),
    # configs[2]: ACM-shaped text + keyphrases
    "acm_2v": dict(D=400_000, K=1000, views=[(100_000, 120, 0.5, 1.0, 1024), (50_000, 6, 0.5, 0.8, 64)]),
    # configs[3]: PubMed-OA-shaped text + MeSH + citations
    "pubmed_3v": dict(D=1_000_000, K=1000, views=[(200_000, 230, 0.5, 1.0, 2048), (28_000, 12, 0.4, 0.9, 128),
                                                  (300_000, 10, 0.9, 0.8, 512)]),
    # configs[4]: 4-view stress, long-tail lengths
    "stress_4v": dict(D=2_000_000, K=2000, views=[(200_000, 200, 1.2, 1.0, 16384), (50_000, 10, 1.0, 0.8, 512),
                                                  (30_000, 8, 1.0, 0.8, 512), (500_000, 6, 1.0, 0.8, 512)]),
    # small shapes for tests
    "tiny_1v": dict(D=300, K=20, views=[(200, 12, 0.6, 1.0, 64)]),
    "tiny_2v": dict(D=300, K=20, views=[(200, 12, 0.6, 1.0, 64), (80, 4, 0.5, 0.8, 16)]),
    "small_1v": dict(D=4000, K=100, views=[(3000, 60, 0.6, 1.0, 512)]),
    "small_3v": dict(D=3000, K=100, views=[(3000, 50, 0.5, 1.0, 512), (500, 8, 0.4, 0.9, 64), (2000, 6, 0.9, 0.8, 128)]),
}


def zipf_base(V, s=1.07):
    w = 1.0 / np.arange(1, V + 1, dtype=np.float64) ** s
    return w / w.sum()


def _topic_word_tables(rng, K, V, support):
    """Per-topic sparse word distributions: `support` words drawn from the Zipf base, Gamma-distributed weights."""
    base = zipf_base(V)
    S = min(V, support)
    cum_base = np.cumsum(base)
    words = np.empty((K, S), dtype=np.int32)
    cums = np.empty((K, S), dtype=np.float64)
    for k in range(K):
        # sampling with replacement from the base then de-duplicating keeps frequent words in most topics
        cand = np.searchsorted(cum_base, rng.random(S * 2), side="right").clip(0, V - 1)
        uniq = np.unique(cand)
        if len(uniq) < S:
            extra = rng.choice(V, S - len(uniq), replace=False)
            uniq = np.unique(np.concatenate([uniq, extra]))
        if len(uniq) < S:   # tiny vocabularies
            uniq = np.resize(uniq, S)
        sel = rng.permutation(uniq)[:S]
        wts = rng.gamma(0.3, 1.0, S) * np.sqrt(base[sel]) + 1e-12
        words[k] = sel
        c = np.cumsum(wts)
        cums[k] = c / c[-1]
    return words, cums


def generate(name_or_cfg, seed=20261018, shard=0, docs=None, support=2048):
    """Returns (K, V list, views list of (doc_off int64[D+1], word_id int32[N])).

    `shard` re-seeds the document stream (same topics/vocabulary, different documents) so that N ranks hold N
    different shards of one corpus (weak scaling); `docs` overrides D."""
    cfg = CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
    K = cfg["K"]
    D = int(docs if docs is not None else cfg["D"])
    topic_rng = np.random.Generator(np.random.Philox(key=seed))
    doc_rng = np.random.Generator(np.random.Philox(key=seed + 7919 * (shard + 1)))
    T = 8                                                     # topics per document (sparse Dir(0.1) stand-in)
    doc_topics = doc_rng.integers(0, K, size=(D, T), dtype=np.int32)
    dw = doc_rng.gamma(0.5, 1.0, size=(D, T)) + 1e-9
    doc_cum = np.cumsum(dw, axis=1)
    doc_cum /= doc_cum[:, -1:]
    views, Vs = [], []
    for m, (V, mean_len, sigma, cover, nmax) in enumerate(cfg["views"]):
        words, cums = _topic_word_tables(topic_rng, K, V, support)
        mu = math.log(mean_len) - 0.5 * sigma * sigma
        lens = np.clip(np.rint(doc_rng.lognormal(mu, sigma, D)), 1, nmax).astype(np.int64)
        if cover < 1.0:
            lens[doc_rng.random(D) >= cover] = 0
        off = np.zeros(D + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        N = int(off[-1])
        word_id = np.empty(N, dtype=np.int32)
        view_topics = doc_rng.integers(0, K, size=(D, 2), dtype=np.int32)   # the 20 % view-specific part
        CH = 4_000_000
        for s in range(0, N, CH):
            e = min(N, s + CH)
            dtok = np.searchsorted(off, np.arange(s, e, dtype=np.int64), side="right") - 1
            u = doc_rng.random(e - s)
            slot = (u[:, None] > doc_cum[dtok]).sum(axis=1).clip(0, T - 1)
            zt = doc_topics[dtok, slot]
            own = doc_rng.random(e - s) < 0.2
            zt = np.where(own, view_topics[dtok, doc_rng.integers(0, 2, e - s)], zt)
            order = np.argsort(zt, kind="stable")
            zs = zt[order]
            bounds = np.searchsorted(zs, np.arange(K + 1))
            out = np.empty(e - s, dtype=np.int32)
            r = doc_rng.random(e - s)
            for k in range(K):
                a, b = bounds[k], bounds[k + 1]
                if a == b:
                    continue
                idx = np.searchsorted(cums[k], r[a:b], side="right").clip(0, cums.shape[1] - 1)
                out[order[a:b]] = words[k][idx]
            word_id[s:e] = out
        views.append((off, word_id))
        Vs.append(V)
    return K, Vs, views


def generate_uniform(D, K, V, mean_len, seed=7):
    """Worst case for the cache hierarchy: words drawn uniformly from V (no Zipf head), fixed-length documents.
    Every token then reads a different n_wk row, so the sweep is genuinely HBM-bound once V*K*4 >> L2."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    off = np.arange(D + 1, dtype=np.int64) * int(mean_len)
    words = rng.integers(0, V, size=int(off[-1]), dtype=np.int32)
    return K, [V], [(off, words)]


def shard_views(views, rank, world):
    """Documents rank, rank+world, ... of every view (all views of a document stay together)."""
    out = []
    for off, word in views:
        D = len(off) - 1
        ids = np.arange(rank, D, world)
        lens = (off[1:] - off[:-1])[ids]
        noff = np.zeros(len(ids) + 1, dtype=np.int64)
        np.cumsum(lens, out=noff[1:])
        idx = np.concatenate([np.arange(off[d], off[d + 1]) for d in ids]) if len(ids) and noff[-1] > 0 else np.zeros(0, dtype=np.int64)
        out.append((noff, word[idx.astype(np.int64)] if len(idx) else np.zeros(0, dtype=np.int32)))
    return out
