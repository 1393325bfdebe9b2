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
