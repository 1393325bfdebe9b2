"""CPU tests: the oracle against the hand-derived known-answer vectors of SURVEY.md section 8(c)
(the reference ships no tests or golden files of its own -> "parity unpinned" beyond these)."""
import json
import os

import numpy as np
import pytest

from helpers import random_corpus, recount

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers(oracle_mod):
    O = oracle_mod
    # Random123 kat_vectors for philox4x32-10
    assert [hex(x) for x in O.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_ftree_known_answers(oracle_mod):
    O = oracle_mod
    # FTree.main (FT:158-161): weights {1,2,3,4}
    t = O.ftree_build([1, 2, 3, 4])
    assert list(t[1:]) == [10, 3, 7, 1, 2, 3, 4]
    assert O.ftree_sample(t, 0.4) == 2
    assert [O.ftree_sample(t, u) for u in (0, 0.0999, 0.1, 0.3, 0.6)] == [0, 0, 1, 2, 3]
    O.ftree_update(t, 0, 5)
    assert t[1] == 14 and t[2] == 7
    # non power-of-two sizes (FT:96-109 works for any K)
    rng = np.random.default_rng(0)
    for K in (3, 5, 50, 500):
        w = rng.random(K)
        t = O.ftree_build(w)
        assert abs(t[1] - w.sum()) < 1e-12
        cnt = np.zeros(K)
        us = rng.random(20000)
        for u in us:
            cnt[O.ftree_sample(t, u)] += 1
        assert np.abs(cnt / len(us) - w / w.sum()).max() < 0.02


def test_lower_bound_known_answers(oracle_mod):
    O = oracle_mod
    assert [O.lower_bound([1, 3, 6, 10], k) for k in (0.5, 1, 1.01, 6, 10, 10.5)] == [0, 0, 1, 2, 3, -1]


def test_log_gamma_stirling_known_answers(oracle_mod):
    O = oracle_mod
    kat = {0.1: 2.2527152267259556, 1: 3.5500763365727898e-06, 2: 3.5500763365727898e-06, 2.1: 0.04544031353623545,
           5: 3.1780538375708223, 100.5: 361.43554046777757}
    for z, v in kat.items():
        assert O.log_gamma_stirling(z) == pytest.approx(v, rel=1e-14, abs=1e-18)


def test_mallet_beta_quirk_q5(oracle_mod):
    O = oracle_mod
    # a > 1, b == 1: first proposal always accepted -> N(1, 0.25/(a-1)) truncated to [0,1], mean well below a/(a+1)
    xs = np.array([O.next_beta_mallet(s, 1.1, 1.0) for s in range(4000)])
    assert xs.min() >= 0 and xs.max() <= 1
    # a < 1: Johnk's algorithm is a true Beta(a,1): E = a/(a+1)
    ys = np.array([O.next_beta_mallet(s, 0.5, 1.0) for s in range(4000)])
    assert abs(ys.mean() - 0.5 / 1.5) < 0.02
    us = np.array([O.next_beta_mallet(s, 1.0, 1.0) for s in range(4000)])
    assert abs(us.mean() - 0.5) < 0.02


def test_histogram_rule(oracle_mod):
    """U:220-232: one doc with n_d[old] 3->2 and n_d[new] 0->1."""
    O = oracle_mod
    K, V = 4, [3]
    off = np.array([0, 3], dtype=np.int64)
    words = np.array([0, 1, 2], dtype=np.int32)
    o = O.Oracle(K, V, [(off, words)], seed=3)
    o.set_assignments([np.array([1, 1, 1], dtype=np.int32)])
    h0 = o.get_hist(0).copy()
    assert h0[1, 3] == 1 and h0[0, 0] == 1
    # sweeps move tokens; the maintained histogram must equal a rebuilt one on bins c >= 1
    for it in range(1, 6):
        o.sweep(it, 0)
        z = o.get_assignments(0)
        h = o.get_hist(0)
        for t in range(K):
            c = int((z == t).sum())
            if c:
                assert h[t, c] == 1
            assert h[t, 1:].sum() == (1 if c else 0)


@pytest.mark.parametrize("flags", [0, 1, 2, 4, 8, 16, 16 | 32, 1 | 2 | 4])
def test_invariants_all_modes(oracle_mod, flags):
    O = oracle_mod
    K, Vs = 17, [40, 25, 9]
    views = random_corpus(5, 150, K, Vs, [9, 4, 2], oov=True)
    o = O.Oracle(K, Vs, views, seed=11)
    o.init_assignments()
    assert o.check_invariants() == 0
    for it in range(1, 8):
        o.sweep(it, flags)
        assert o.check_invariants() == 0
    zs = [o.get_assignments(m) for m in range(3)]
    for m, (nwk, nk) in enumerate(recount(views, zs, K, Vs)):
        a, b = o.get_counts(m)
        assert np.array_equal(a, nwk) and np.array_equal(b, nk)
        assert b.sum() == len(zs[m])


def test_mt_scheme_consistent_at_barrier(oracle_mod):
    O = oracle_mod
    K, Vs = 20, [60, 30]
    views = random_corpus(7, 400, K, Vs, [12, 3])
    o = O.Oracle(K, Vs, views, seed=2)
    o.init_assignments()
    for it in range(1, 6):
        o.sweep_mt(it, 8)
        assert o.check_invariants() == 0      # a12: at the barrier every delta is applied


def test_net_distribution_identity(oracle_mod):
    """the three buckets sum to phi_t*(p_mm*n_d + [t in S]*O + gamma*alpha) (+C)."""
    O = oracle_mod
    K, Vs = 12, [30, 20]
    views = random_corpus(9, 60, K, Vs, [8, 5], empty_frac=0.0)
    o = O.Oracle(K, Vs, views, seed=4)
    rng = np.random.default_rng(1)
    alpha = rng.uniform(0.01, 0.3, size=(2, K + 1))
    o.set_hyper(alpha=alpha, alphaSum=alpha.sum(1), gamma=[0.7, 1.9], inactive=[3, 5])
    o.init_assignments()
    z0 = o.get_assignments(0); z1 = o.get_assignments(1)
    z0[(z0 == 3) | (z0 == 5)] = 1; z1[(z1 == 3) | (z1 == 5)] = 2
    o.set_assignments([z0, z1])
    p = [[1.0, 0.37], [0.37, 1.0]]
    for d in range(10):
        for m in range(2):
            a = o.cond_probs(m, d, 0, p=p)
            b = o.cond_probs(m, d, 0, p=p, engine_form=True)
            assert np.allclose(a, b, rtol=1e-12, atol=0)
            assert abs(a[:K].sum() - 1) < 1e-12 and a[K] > 0


def test_loglik_matches_scipy_gammaln(oracle_mod):
    from scipy.special import gammaln
    O = oracle_mod
    K, Vs = 10, [30]
    views = random_corpus(3, 80, K, Vs, [10], empty_frac=0.2)
    o = O.Oracle(K, Vs, views, seed=5)
    o.init_assignments()
    z = o.get_assignments(0)
    off = views[0][0]
    nwk, nk = o.get_counts(0)
    a, b, g = 0.1, 0.01, 1.0
    ll = 0.0
    ndocs = 0
    for d in range(len(off) - 1):
        zz = z[off[d]:off[d + 1]]
        if len(zz) == 0:
            continue
        c = np.bincount(zz, minlength=K)
        ll += (gammaln(g * a + c[c > 0]) - gammaln(g * a)).sum() - gammaln(g * a * K + len(zz))
        ndocs += 1
    ll += ndocs * gammaln(g * a * K)
    ll += gammaln(b + nwk[nwk > 0]).sum() - gammaln(b * Vs[0] + nk).sum() + K * gammaln(b * Vs[0]) - (nwk > 0).sum() * gammaln(b)
    got = o.loglik()[0]
    # logGammaStirling is only good to ~3.6e-6 per term
    assert got == pytest.approx(ll, rel=1e-5)
    # Q18: phantom topic-0 tokens make the quirk value differ when short documents exist
    assert o.loglik(quirk_len2=True)[0] != got


def test_golden_fixture(oracle_mod):
    """Committed fixture generated by tests/golden/make_golden.py from this oracle: guards against drift."""
    O = oracle_mod
    path = os.path.join(GOLDEN, "oracle_tiny.json")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated yet")
    g = json.load(open(path))
    views = [(np.array(v["off"], dtype=np.int64), np.array(v["word"], dtype=np.int32)) for v in g["views"]]
    o = O.Oracle(g["K"], g["V"], views, seed=g["seed"])
    o.init_assignments()
    for m in range(len(views)):
        assert o.get_assignments(m).tolist() == g["z_init"][m]
    for it in range(1, g["sweeps"] + 1):
        o.sweep(it, g["flags"])
    for m in range(len(views)):
        assert o.get_assignments(m).tolist() == g["z_final"][m]
    assert np.allclose(o.loglik(), g["loglik"], rtol=1e-12)
