"""CPU restatement (numpy / pure Python, small cases) of the reference's hyper-parameter step -- TEST INFRASTRUCTURE.

Follows org.madgik.MVTopicModel.FastQMVWVParallelTopicModel (M): optimizeP M:2698-2819, optimizeBeta M:2288-2367,
the Antoniak law behind optimizeDP (org.knowceans.util.Samplers.stirling / randAntoniak, KS:1052-1110) and MALLET 2.0.8's
Dirichlet.digamma / learnSymmetricConcentration as recovered in SURVEY.md section 8(c) (binary-only dependency).
PARITY STATUS: pinned to outputs of the reference's own binaries, executed by tools/jvm_mini.py (the reference has no tests of its
own): optimizeBeta and MALLET's digamma / learnSymmetricConcentration (tests/test_reference_vectors.py), optimizeP's per-pair sums,
pMean and p_a on corpora where the older jar and the source agree (tests/test_optim_host.py; the TreeMap collision rule Q11 rests
on the source citation), the Antoniak law against a first call of Samplers.randAntoniak.  These restatements are what the engine's
device statistics and host code in mvtopicmodel_b200/csrc/mvtm_optim.inl are checked against.
"""
import math

import numpy as np


def p_statistics(views, zs, K, present=None):
    """Sum over documents of pDistr_Mean[m][i][doc] (M:2706-2782) with the TreeMap collision rule (Q11)."""
    M = len(views)
    D = len(views[0][0]) - 1
    psum = np.zeros((M, M))
    for d in range(D):
        lens = [int(views[m][0][d + 1] - views[m][0][d]) for m in range(M)]
        tm = {}
        for m in range(M):
            tm[lens[m]] = m                          # sortedViews.put(length, m): equal lengths overwrite
        order = [tm[k] for k in sorted(tm, reverse=True)]
        topics = [set(int(t) for t in zs[m][views[m][0][d]:views[m][0][d + 1]] if t >= 0) for m in range(M)]
        prev = [order[0]]
        for m in order[1:]:
            if lens[m] > 0:
                zz = zs[m][views[m][0][d]:views[m][0][d + 1]]
                for i in prev:
                    v = sum(1.0 for t in zz if t >= 0 and int(t) in topics[i]) / lens[m]
                    psum[m, i] += v
                    psum[i, m] += v
            prev.append(m)
    return psum


def p_params(psum, docs_per_view):
    M = psum.shape[0]
    pa, pmean = np.full((M, M), 0.2), np.eye(M)
    for m in range(M):
        for i in range(m + 1, M):
            mean = psum[m, i] / min(docs_per_view[m], docs_per_view[i])
            a = 5000.0 if mean == 1 else (-1.0 / math.log(mean) if mean > 0 else 0.0)
            pa[m, i] = pa[i, m] = min(a, 100.0)
            pmean[m, i] = pmean[i, m] = mean
    return pa, pmean


def _jdiv(a, b):
    """IEEE-754 division as the JVM performs it (x/0 = +-Infinity, 0/0 = NaN) -- Python raises instead."""
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def mallet_digamma(z):
    if z != z:
        return math.nan
    if z < 1e-6:
        return -0.5772156649015329 - _jdiv(1.0, z)
    acc = 0.0
    while z < 9.5:
        acc -= _jdiv(1.0, z)
        z += 1.0
    return acc + (math.log(z) if z > 0 and z != math.inf else (math.inf if z == math.inf else math.nan)) - _jdiv(1.0, 2.0 * z)


def learn_symmetric_concentration(count_hist, length_hist, num_dims, current):
    largest = max([i for i, c in enumerate(count_hist) if c > 0], default=0)
    nz = [i for i, c in enumerate(length_hist) if c > 0]
    for _ in range(200):
        param = current / num_dims
        dg, num = 0.0, 0.0
        for idx in range(1, largest + 1):
            dg += _jdiv(1.0, param + idx - 1)
            num += count_hist[idx] * dg
        dg, den = 0.0, 0.0
        cached = mallet_digamma(current)
        for length in nz:
            if length > 20:          # previousLength never advances in MALLET 2.0.8
                dg = mallet_digamma(current + length) - cached
            else:
                for idx in range(0, length):
                    dg += _jdiv(1.0, current + idx)
            den += dg * length_hist[length]
        current = _jdiv(param * num, den)
    return current


def optimize_beta(nwk, nk, V, beta, beta_sum):
    vals = nwk[nwk > 0]
    count_hist = np.bincount(vals, minlength=int(nk.max()) + 1)
    size_hist = np.bincount(nk, minlength=int(nk.max()) + 1)
    bs = learn_symmetric_concentration(count_hist.tolist(), size_hist.tolist(), V, beta_sum)
    if bs < V * 0.0001:
        return 0.0001, 0.0001 * V
    if math.isnan(bs):
        return (0.0001, 0.0001 * V) if beta == 0.01 else (beta_sum / V, beta_sum)
    return bs / V, bs


def antoniak_pmf(alpha, n):
    """P(number of tables = m), m = 1..n, by the reference's normalised Stirling recurrence (KS:1052-1078)."""
    ss = np.array([1.0])
    for mm in range(1, n):
        new = np.zeros(len(ss) + 1)
        new[:-1] += ss * mm
        new[1:] += ss
        ss = new / new.max()
    p = ss * alpha ** np.arange(len(ss))
    # rescale in log space for large n / alpha
    if not np.all(np.isfinite(p)) or p.sum() == 0:
        lp = np.log(np.maximum(ss, 1e-300)) + np.arange(len(ss)) * math.log(alpha)
        p = np.exp(lp - lp.max())
    return p / p.sum()
