"""CPU tests of the host-side hyper-parameter code in libmvtm.so (no device needed): the samplers against their exact
laws, MALLET's learnSymmetricConcentration against the oracle restatement."""
import ctypes as C

import numpy as np
import pytest


def draws(lib, seed, which, a, b, n):
    out = np.empty(n, dtype=np.float64)
    assert lib.mvtm_test_sampler(int(seed), int(which), float(a), float(b), int(n), out.ctypes.data_as(C.c_void_p)) == 0
    return out


def test_uniform_stream_is_philox(engine_lib, oracle_mod):
    u = draws(engine_lib, 0x1234567890ABCDEF, 0, 0, 0, 4)
    for n, got in enumerate(u):
        x = oracle_mod.philox([n, 0, 0, 3], [0x90ABCDEF, 0x12345678])
        want = float(((int(x[0]) >> 5) << 26) | (int(x[1]) >> 6)) / 2.0**53
        assert got == want


@pytest.mark.parametrize("shape", [0.05, 0.5, 1.0, 2.5, 40.0, 900.0])
def test_gamma_sampler_moments(engine_lib, shape):
    x = draws(engine_lib, 7, 1, shape, 0, 200_000)
    assert x.min() >= 0
    assert x.mean() == pytest.approx(shape, rel=0.02)
    assert x.var() == pytest.approx(shape, rel=0.05)


@pytest.mark.parametrize("a,b", [(1.0, 1.0), (2.0, 5.0), (0.3, 1.0), (11.0, 250.0)])
def test_beta_sampler_moments(engine_lib, a, b):
    x = draws(engine_lib, 9, 2, a, b, 200_000)
    assert 0 <= x.min() and x.max() <= 1
    assert x.mean() == pytest.approx(a / (a + b), rel=0.02)
    assert x.var() == pytest.approx(a * b / ((a + b) ** 2 * (a + b + 1)), rel=0.06)
    # KR:267-271 with a zero parameter: randGamma(0) = 0 -> Beta(a, 0) = 1
    assert np.all(draws(engine_lib, 9, 2, 1.5, 0.0, 10) == 1.0)


@pytest.mark.parametrize("alpha,n", [(0.1, 2), (0.1, 7), (1.0, 30), (3.7, 12), (0.02, 200)])
def test_antoniak_sampler_matches_stirling_law(engine_lib, alpha, n):
    """The engine samples the number of tables as a sum of Bernoulli(alpha/(alpha+i)); the reference inverts the Stirling
    table (KS:1089-1110).  Same law: compare the empirical pmf with the exact one (chi-square style bound)."""
    from oracle import optim
    N = 200_000
    x = draws(engine_lib, 11, 3, alpha, n, N).astype(np.int64)
    pmf = optim.antoniak_pmf(alpha, n)
    emp = np.bincount(x, minlength=n + 1)[1:n + 1] / N
    assert np.abs(emp - pmf).max() < 5 * np.sqrt(pmf.max() / N) + 1e-4
    assert x.mean() == pytest.approx(sum(alpha / (alpha + i) for i in range(n)), rel=0.02, abs=0.01)
    # MAXSTIRLING (KS:1023): beyond 20000 the reference throws and optimizeDP falls back to one table (M:2470-2475)
    assert np.all(draws(engine_lib, 1, 3, 0.5, 20001, 3) == 1)


def test_learn_symmetric_concentration_matches_oracle(engine_lib):
    from oracle import optim
    rng = np.random.default_rng(0)
    for V, K in [(300, 20), (5000, 100)]:
        nk = rng.integers(50, 4000, size=K)
        vals = rng.geometric(0.3, size=K * 60)
        count_hist = np.bincount(vals, minlength=int(nk.max()) + 1).astype(np.int64)
        size_hist = np.bincount(nk, minlength=int(nk.max()) + 1).astype(np.int64)
        for cur in (0.01 * V, 0.5 * V):
            got = engine_lib.mvtm_test_learn_symmetric_concentration(count_hist.ctypes.data_as(C.c_void_p), len(count_hist),
                                                                     size_hist.ctypes.data_as(C.c_void_p), len(size_hist), V, cur)
            want = optim.learn_symmetric_concentration(count_hist.tolist(), size_hist.tolist(), V, cur)
            assert got == pytest.approx(want, rel=1e-12)
            assert got > 0


def test_mallet_digamma_quirk_q19():
    """With the Bernoulli terms folded to 0 the value differs from the true digamma by about 1/(12 z^2) at z ~ 10."""
    from oracle import optim
    from scipy.special import digamma
    assert optim.mallet_digamma(0.5) != pytest.approx(float(digamma(0.5)), abs=1e-6)
    assert optim.mallet_digamma(0.5) == pytest.approx(float(digamma(0.5)), abs=2e-3)
    assert optim.mallet_digamma(1e-7) == pytest.approx(-0.5772156649015329 - 1e7)


# ---- the same samplers against draws made by the reference's own sampler bytecode (tests/golden/make_reference_hyper_vectors.py) ----

def _hyper_vectors():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_hyper_vectors.json")))


def _ks_two_sample(x, y):
    x, y = np.sort(x), np.sort(y)
    grid = np.concatenate([x, y])
    d = np.abs(np.searchsorted(x, grid, side="right") / len(x) - np.searchsorted(y, grid, side="right") / len(y)).max()
    return d, 1.95 * np.sqrt((len(x) + len(y)) / (len(x) * len(y)))          # Kolmogorov-Smirnov bound at alpha = 0.001


def test_gamma_sampler_matches_reference_bytecode_law(engine_lib):
    """randGamma(shape, scale) of the reference (KR:294-360) against rand_gamma(shape) * scale of the engine: same law, and the
    second argument is a SCALE (optimizeGamma passes 1/bloge, M:2394)."""
    for k, rec in enumerate(_hyper_vectors()["randGamma"]):
        ref = np.array(rec["samples"])
        got = draws(engine_lib, 100 + k, 1, rec["shape"], 0, 30_000) * rec["scale"]
        d, bound = _ks_two_sample(ref, got)
        assert d < bound, (rec["shape"], rec["scale"], d, bound)
        assert ref.mean() == pytest.approx(rec["shape"] * rec["scale"], rel=0.15)


def test_gamma_sampler_matches_mallet_bytecode_law(engine_lib):
    """Randoms.nextGamma(alpha, 1) of the MALLET jar -- what sampleDirichlet draws (M:2616) -- against rand_gamma of the engine."""
    recs = _hyper_vectors()["mallet_nextGamma"]
    assert len(recs) >= 6
    for k, rec in enumerate(recs):
        d, bound = _ks_two_sample(np.array(rec["samples"]), draws(engine_lib, 400 + k, 1, rec["shape"], 0, 30_000))
        assert d < bound, (rec["shape"], d, bound)


def test_beta_sampler_matches_reference_bytecode_law(engine_lib):
    for k, rec in enumerate(_hyper_vectors()["randBeta"]):
        ref = np.array(rec["samples"])
        got = draws(engine_lib, 200 + k, 2, rec["a"], rec["b"], 30_000)
        d, bound = _ks_two_sample(ref, got)
        assert d < bound, (rec["a"], rec["b"], d, bound)


def test_bernoulli_of_reference_bytecode():
    """randBernoulli(p) is 1 with probability p (KR:789-795), the form the engine uses (u < p)."""
    for rec in _hyper_vectors()["randBernoulli"]:
        p, n = rec["p"], rec["n"]
        assert abs(rec["ones"] / n - p) <= 4 * np.sqrt(max(p * (1 - p), 1e-12) / n)


def test_antoniak_first_call_of_reference_bytecode_is_the_stirling_law(engine_lib):
    """A first randAntoniak(alpha, n) on a fresh class inverts exactly the law the engine samples (KS:1089-1110): the table count it
    returns for a uniform u is the bin of the exact CDF holding u.  Repeated calls drift away from it (Q7: the cached Stirling row is
    multiplied and prefix-summed in place) -- the engine keeps the exact law."""
    from oracle import optim
    hv = _hyper_vectors()
    for rec in hv["randAntoniak_first_call"]:
        cdf = np.cumsum(optim.antoniak_pmf(rec["alpha"], rec["n"]))
        for u, tables in rec["draws"]:
            lo = cdf[tables - 2] if tables >= 2 else 0.0
            assert lo - 1e-12 <= u <= cdf[tables - 1] + 1e-12, (rec["alpha"], rec["n"], u, tables)
    drift = []
    for rec in hv["randAntoniak_repeated"]:
        x = np.array(rec["draws"], dtype=np.float64)
        drift.append(abs(x.mean() - rec["exact_mean"]) / (x.std() / np.sqrt(len(x))))   # the reference's own drift, in standard errors
        got = draws(engine_lib, 300, 3, rec["alpha"], rec["n"], 100_000)
        assert got.mean() == pytest.approx(rec["exact_mean"], rel=0.01)                 # the engine: the exact law
    assert max(drift) > 10                                                              # (3.7, 12): mean 7.3 against 5.75


# ---- optimizeDP / optimizeGamma as a whole: the product's host code against the reference's bytecode fed the same draws ----

def _run_hyper_core(lib, case, which, script, state=None):
    """mvtm_test_hyper_core on a case of reference_hyper_step_vectors.json; returns the state dict after the call and the arg log"""
    M, K = case["M"], case["K"]
    hist = [np.ascontiguousarray(h, dtype=np.int64) for h in case["topicDocCounts"]]
    lencnt = [np.ascontiguousarray(l, dtype=np.int64) for l in case["docLengthCounts"]]
    stride = np.array([h.shape[1] for h in hist], dtype=np.int32)
    n_len = np.array([len(l) for l in lencnt], dtype=np.int32)
    hp = (C.c_void_p * M)(*[h.ctypes.data for h in hist])
    lp = (C.c_void_p * M)(*[l.ctypes.data for l in lencnt])
    st = state or {"alpha": np.array(case["in"]["alpha"], dtype=np.float64), "alphaSum": np.array([sum(a) for a in case["in"]["alpha"]]),
                   "gamma": np.array(case["in"]["gamma"]), "gammaView": np.array(case["in"]["gammaView"]), "tablesCnt": np.zeros(M),
                   "scal": np.array([case["in"]["gammaRoot"], 0.0]), "inactive": []}
    vals = np.ascontiguousarray([s[3] for s in script], dtype=np.float64)
    log = np.zeros((max(len(script), 1), 3))
    inact, n_inact, n_used = np.zeros(K, dtype=np.int32), C.c_int32(-1), C.c_int64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.mvtm_test_hyper_core(M, K, which, hp, p(stride), lp, p(n_len), p(st["alpha"]), p(st["alphaSum"]), p(st["gamma"]),
                                  p(st["gammaView"]), p(st["tablesCnt"]), p(st["scal"]), p(inact), C.byref(n_inact),
                                  p(vals), len(vals), p(log), C.byref(n_used))
    if n_inact.value >= 0:
        st["inactive"] = inact[:n_inact.value].tolist()
    return rc, st, log, n_used.value


def _check_draw_arguments(script, log):
    want = np.array([s[:3] for s in script], dtype=np.float64)
    assert np.array_equal(log[:, 0], want[:, 0])                       # the same sampler at every step of the script
    np.testing.assert_allclose(log[:, 1], want[:, 1], rtol=1e-12, atol=0)
    two_arg = want[:, 0] != 1                                          # Gamma: the engine applies the scale itself (checked through the results)
    np.testing.assert_allclose(log[two_arg, 2], want[two_arg, 2], rtol=1e-12, atol=0)


def test_optimize_dp_and_gamma_match_reference_bytecode_on_scripted_draws(engine_lib):
    """The functions behind mvtm_optimize_hyper(MVTM_OPT_DP | MVTM_OPT_GAMMA), fed the values the reference's optimizeDP (M:2440-2591)
    and optimizeGamma (M:2369-2438) bytecode received from its samplers: same draws requested, in the same order, with the same
    arguments (1e-12), and the same hyper-parameters at the end.  tests/golden/make_reference_hyper_step_vectors.py."""
    import json
    import os
    from mvtopicmodel_b200 import _lib
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_hyper_step_vectors.json")))["cases"]
    assert len(cases) >= 3
    for case in cases:
        script, n_dp = case["script"], case["after_optimizeDP"]["draws"]
        # optimizeDP alone
        rc, st, log, used = _run_hyper_core(engine_lib, case, _lib.OPT_DP, script[:n_dp])
        assert rc == 0 and used == n_dp, (case["name"], rc, used, n_dp)
        _check_draw_arguments(script[:n_dp], log)
        ref = case["after_optimizeDP"]
        np.testing.assert_allclose(st["alpha"], np.array(ref["alpha"]), rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(st["alphaSum"], ref["alphaSum"], rtol=1e-12)
        np.testing.assert_allclose(st["tablesCnt"], ref["tablesCnt"], rtol=1e-12)
        assert st["scal"][1] == pytest.approx(ref["rootTablesCnt"], rel=1e-12)
        assert st["inactive"] == ref["inactive"]
        assert abs(st["alpha"].sum(axis=1) - 1).max() < 1e-9
        # ... then optimizeGamma on that state (the chain of M:1186-1193)
        rc, st, log, used = _run_hyper_core(engine_lib, case, _lib.OPT_GAMMA, script[n_dp:], state=st)
        assert rc == 0 and used == len(script) - n_dp, (case["name"], rc, used)
        _check_draw_arguments(script[n_dp:], log)
        ref = case["after_optimizeGamma"]
        assert st["scal"][0] == pytest.approx(ref["gammaRoot"], rel=1e-12)
        np.testing.assert_allclose(st["gammaView"], ref["gammaView"], rtol=1e-12)
        np.testing.assert_allclose(st["gamma"], ref["gamma"], rtol=1e-12)
        # both in one call, and a script that is too short is reported, not silently padded
        rc, st2, _, used = _run_hyper_core(engine_lib, case, _lib.OPT_DP | _lib.OPT_GAMMA, script)
        assert rc == 0 and used == len(script)
        np.testing.assert_allclose(st2["gamma"], ref["gamma"], rtol=1e-12)
        rc, _, _, _ = _run_hyper_core(engine_lib, case, _lib.OPT_DP, script[:n_dp - 1])
        assert rc == 1


def test_optimize_p_restatement_matches_reference_bytecode():
    """oracle/optim.py p_statistics + p_params (what k_p_stats and the engine's optimize_p are held to on the GPU, 1e-12) vs
    optimizeP executed from the reference's jar (tests/golden/make_reference_optimize_p_vectors.py): per-pair sums of
    pDistr_Mean on every case; pMean and p_a = min(-1/ln pMean, 100) incl. the pMean = 1 -> 5000 -> 100 branch where every document
    has every view (the older jar divides by totalDocsPerModality[m] instead of the min over the pair, M:2793)."""
    import json
    import os
    from oracle import optim
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_optimize_p.json")))
    assert len(g["cases"]) >= 4
    saw_cap = False
    for c in g["cases"]:
        views = [(np.array(v["off"]), np.array(v["word"])) for v in c["views"]]
        zs = [np.array(z) for z in c["z"]]
        ps = optim.p_statistics(views, zs, c["K"])
        docs = np.array(c["totalDocsPerModality"], dtype=np.float64)
        iu = np.triu_indices(c["M"], 1)
        np.testing.assert_allclose(ps[iu], (np.array(c["pMean"]) * docs[:, None])[iu], rtol=1e-12)
        assert np.array_equal(ps, ps.T)
        if c["all_views_present"]:
            pa, pm = optim.p_params(ps, c["totalDocsPerModality"])
            off = ~np.eye(c["M"], dtype=bool)
            np.testing.assert_allclose(pm, np.array(c["pMean"]), rtol=1e-12)
            np.testing.assert_allclose(pa[off], np.array(c["p_a"])[off], rtol=1e-12)
            assert np.all(np.array(c["p_b"])[off] == 1.0)
            saw_cap |= bool(np.any(pa[off] == 100.0))
    assert saw_cap


def test_engine_mallet_beta_law_matches_mallet_bytecode(engine_lib):
    """MVTM_FLAG_BETA_MALLET: the sweep kernel's view-coupling draw (mallet_next_beta in mvtm_kernels.cuh, a __host__ __device__
    function evaluated here on the host through mvtm_test_sampler which = 4) against draws of cc.mallet.util.Randoms.nextBeta
    from the MALLET jar's own bytecode (tests/golden/reference_beta_vectors.json): two-sample Kolmogorov-Smirnov at
    alpha = 0.001 on every (a, b) pair, incl. quirk Q5 -- for a > 1, b = 1 a truncated normal whose mean lies below a/(a+1)."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_beta_vectors.json")))
    assert len(g["cases"]) >= 8
    for k, case in enumerate(g["cases"]):
        ref = np.sort(np.array(case["samples"]))
        got = np.sort(draws(engine_lib, 1000003 * k + 17, 4, case["a"], case["b"], 6000))
        assert 0.0 <= got[0] and got[-1] <= 1.0
        grid = np.concatenate([ref, got])
        d = np.abs(np.searchsorted(ref, grid, side="right") / len(ref) - np.searchsorted(got, grid, side="right") / len(got)).max()
        assert d < 1.95 * np.sqrt((len(ref) + len(got)) / (len(ref) * len(got))), (case["a"], case["b"], d)
        if case["a"] >= 2 and case["b"] == 1:
            assert got.mean() < case["a"] / (case["a"] + 1) - 0.01      # Q5: not the Beta(a, 1) law
