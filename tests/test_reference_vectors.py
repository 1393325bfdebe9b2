"""Pins against outputs of THE REFERENCE'S OWN BINARIES (tests/golden/reference_vectors.json, produced by executing the bytecode
of output/MVTopicModel-1.0-SNAPSHOT.jar and output/lib/mallet-2.0.8.jar with tools/jvm_mini.py -- there is no JVM in this image):
the C oracle's F+tree and lower_bound, the Stirling log-gamma used by the log-likelihood, MALLET's digamma and
learnSymmetricConcentration as restated in oracle/optim.py AND in the product's host code (libmvtm.so)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def test_ftree_matches_reference_bytecode(oracle_mod):
    """org.madgik.utils.FTree (FT:96-147): constructor, sample(u), update(t, v) -- tree arrays bit for bit, every draw identical."""
    O = oracle_mod
    for rec in GOLD["ftree"]:
        K = rec["K"]
        tree = O.ftree_build(rec["weights"])
        assert np.array_equal(tree[1:], np.array(rec["tree"])[1:]), K          # slot 0 is unused by both
        for u, want in rec["samples"]:
            assert O.ftree_sample(tree, u) == want, (K, u)
        for up in rec["updates"]:
            O.ftree_update(tree, up["topic"], up["value"])
            assert np.array_equal(tree[1:], np.array(up["tree"])[1:]), (K, up["topic"])
            for u, want in up["samples"]:
                assert O.ftree_sample(tree, u) == want, (K, u)


def test_lower_bound_matches_reference_bytecode(oracle_mod):
    """FastQMVWVWorkerRunnable.lower_bound (W:257-277; the prebuilt jar's older build takes int[]: same search)."""
    for rec in GOLD["lower_bound"]:
        assert oracle_mod.lower_bound([float(x) for x in rec["arr"]], rec["key"], rec["n"]) == rec["result"], rec


def test_log_gamma_stirling_matches_mallet_bytecode(oracle_mod, engine_lib):
    """cc.mallet.types.Dirichlet.logGammaStirling, the lgamma of modelLogLikelihood (M:3343 ...)."""
    for z, want in GOLD["logGammaStirling"]:
        got = oracle_mod.log_gamma_stirling(z)
        assert got == pytest.approx(want, rel=1e-15, abs=1e-300), (z, got, want)


def test_digamma_and_concentration_match_mallet_bytecode(engine_lib):
    """cc.mallet.types.Dirichlet.digamma (the Bernoulli-free series of Q19) and learnSymmetricConcentration (optimizeBeta,
    M:2327): the numpy restatement and the PRODUCT's host implementation against the jar's own results."""
    from oracle import optim
    for z, want in GOLD["digamma"]:
        assert optim.mallet_digamma(z) == pytest.approx(want, rel=1e-14), z
    f = engine_lib.mvtm_test_learn_symmetric_concentration
    for rec in GOLD["learnSymmetricConcentration"]:
        ch = np.asarray(rec["countHistogram"], dtype=np.int64); sh = np.asarray(rec["topicSizeHistogram"], dtype=np.int64)
        want = rec["result"]
        got_np = optim.learn_symmetric_concentration(ch.tolist(), sh.tolist(), rec["numDimensions"], rec["currentValue"])
        got_lib = f(ch.ctypes.data_as(C.c_void_p), len(ch), sh.ctypes.data_as(C.c_void_p), len(sh), rec["numDimensions"], rec["currentValue"])
        # half of the cases end in NaN -- the outcome optimizeBeta's `Double.isNaN(betaSum)` branch exists for (M:2340-2344)
        assert got_np == pytest.approx(want, rel=1e-12, nan_ok=True), rec["numDimensions"]
        assert got_lib == pytest.approx(want, rel=1e-12, nan_ok=True), rec["numDimensions"]


def test_sampler_matches_reference_bytecode(oracle_mod):
    """THE SAMPLER against the reference's own binary: tests/golden/reference_sampler_vectors.json holds assignments produced by
    executing FastQMVWVWorkerRunnable.sampleTopicsForOneDoc (W:301-597) from output/MVTopicModel-1.0-SNAPSHOT.jar, document by
    document and sweep by sweep, with the uniforms the oracle draws for the same token (make_reference_sampler_vectors.py).  The C
    oracle in reference-faithful mode (stale F+trees maintained per delta U:242-260, dead insertion code Q1, deltas applied at
    once) must reproduce them TOKEN FOR TOKEN: single view; two coupled views (Beta-drawn p, other-view mass W:399-410); three
    views with inactive topics (new-topic bucket W:413-418 / W:515, activation U:263-270).  The deltas were applied by the
    jar's FastQMVWVUpdaterRunnable.run, whose doc-topic histogram must equal the oracle's at the end.  After every sweep the oracle's
    log-likelihood must equal the value the jar's modelLogLikelihood bytecode returns for that state."""
    O = oracle_mod
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_sampler_vectors.json")))
    for case in gold["cases"]:
        K, Vs = case["K"], case["V"]
        views = [(np.array(v["off"], dtype=np.int64), np.array(v["word"], dtype=np.int32)) for v in case["views"]]
        # the reference keeps an Assignments[0] object for every document and one for view m > 0 only where the document has
        # the view (M:437-455); the log-likelihood's phantom-token quirk Q18 looks at exactly that
        present = [np.ones(len(views[0][0]) - 1, dtype=np.uint8)] + [((v[0][1:] - v[0][:-1]) > 0).astype(np.uint8) for v in views[1:]]
        o = O.Oracle(K, Vs, views, seed=case["seed"], present=present)
        o.set_hyper(alpha=np.array(case["alpha"]), alphaSum=np.array(case["alphaSum"]), beta=np.array(case["beta"]),
                    betaSum=np.array(case["betaSum"]), gamma=np.array(case["gamma"]), p_a=np.array(case["p_a"]), p_b=np.array(case["p_b"]), inactive=case["inactive"])
        o.set_assignments([np.array(z, dtype=np.int32) for z in case["z0"]])
        o.rebuild_trees()
        total = mism = 0
        for it, want in enumerate(case["z_after"], start=1):
            o.sweep(it, O.F_STALE_TREES | O.F_Q1_COMPAT)
            for m in range(len(Vs)):
                got = o.get_assignments(m)
                w = np.array(want[m], dtype=np.int32)
                total += len(w); mism += int((got != w).sum())
                assert np.array_equal(got, w), (case["name"], "sweep", it, "view", m, "first mismatch at", int(np.argmax(got != w)))
            # modelLogLikelihood (M:3322-3452) executed from the jar on the same state, incl. the phantom tokens of Q18
            assert np.allclose(o.loglik(True), case["loglik_after"][it - 1], rtol=1e-13, atol=0), (case["name"], it)
        for m in range(len(Vs)):
            assert o.get_counts(m)[1].tolist() == case["nk_final"][m], case["name"]
            # topicDocCounts as the jar's UPDATER left it after all those deltas (U:220-232) vs the oracle's maintained copy
            want, got = np.array(case["hist_maintained"][m]), o.get_hist(m)
            w = min(want.shape[1], got.shape[1])
            assert np.array_equal(got[:, 1:w], want[:, 1:w]) and not want[:, w:].any() and not got[:, w:].any(), (case["name"], m)
        assert total > 0 and mism == 0


def _conditional_cases():
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_sampler_vectors.json")))
    for case in gold["cases"]:
        K, Vs = case["K"], case["V"]
        views = [(np.array(v["off"], dtype=np.int64), np.array(v["word"], dtype=np.int32)) for v in case["views"]]
        yield case, K, Vs, views


def _state_with_doc(case, views, rec):
    """The frozen-count state's assignments with document rec['doc'] as the reference held it when it reached the token."""
    zs = [np.array(z, dtype=np.int32) for z in case["frozen_counts_z"]]
    for m, zd in enumerate(rec["z_doc"]):
        if zd is not None:
            b = int(views[m][0][rec["doc"]])
            zs[m][b:b + len(zd)] = zd
    return zs


def test_conditionals_match_reference_bytecode(oracle_mod):
    """north_star check (b) for the ORACLE: per-token conditional distributions on frozen counts as the reference's sampler
    bytecode computed them (its S / cumulative masses / C / F+tree leaves read out of the running frame) vs the oracle's
    three-bucket masses and its engine-form dense distribution: 1e-9 relative / 1e-15 absolute.  Tokens on which quirk Q1 has
    had an effect (topics gained earlier in the sweep are missing from the reference's dense index) are included: the vectors
    name those topics (`not_in_S`) and both forms must drop their document terms exactly as the reference does."""
    O = oracle_mod
    n = n_q1 = 0
    for case, K, Vs, views in _conditional_cases():
        M = len(Vs)
        o = O.Oracle(K, Vs, views, seed=case["seed"])
        o.set_hyper(alpha=np.array(case["frozen_alpha"]), alphaSum=np.array(case["alphaSum"]), beta=np.array(case["beta"]),
                    betaSum=np.array(case["betaSum"]), gamma=np.array(case["gamma"]), inactive=case["frozen_inactive"])
        o.set_assignments([np.array(z, dtype=np.int32) for z in case["frozen_counts_z"]])
        frozen = [o.get_counts(m) for m in range(M)]
        for rec in case["conditionals"][::3] + [r for r in case["conditionals"] if r.get("not_in_S")][::2]:
            o.set_assignments(_state_with_doc(case, views, rec))
            for m in range(M):
                o.set_counts(m, *frozen[m])
            o.rebuild_trees()
            p = np.eye(M); p[rec["view"]] = rec["p_row"]
            for engine_form in (False, True):       # the three-bucket masses and the engine's dense net distribution
                got = o.cond_probs(rec["view"], rec["doc"], rec["pos"], p=p, engine_form=engine_form, not_in_S=rec.get("not_in_S"))
                # (the reference's document masses are recovered as differences of its cumulative array: absolute error ~1e-16)
                assert np.allclose(got[:K], rec["probs"], rtol=1e-9, atol=1e-15), (case["name"], rec["doc"], rec["view"], rec["pos"], engine_form)
                assert got[K] == pytest.approx(rec["new_share"], rel=1e-9, abs=1e-15)
            n += 1
            n_q1 += bool(rec.get("not_in_S"))
    assert n > 200 and n_q1 > 100      # incl. tokens on which quirk Q1 (dead insertion code) changes the distribution


def test_counts_histograms_and_optimize_beta_match_reference_bytecode(oracle_mod):
    """buildInitialTypeTopicCounts + initializeHistograms (M:600-652, M:849-897) and optimizeBeta (M:2288-2367, through MALLET's
    learnSymmetricConcentration) executed from the jars on the states the reference's sampler reached, vs the oracle: count
    tables and topicDocCounts (bins c >= 1: bin 0 is never read, M:2461) bit for bit; beta / betaSum incl. the 'too sparse'
    sentinel and NaN branches to 1e-12."""
    from oracle import optim
    O = oracle_mod
    for case, K, Vs, views in _conditional_cases():
        M = len(Vs)
        o = O.Oracle(K, Vs, views, seed=case["seed"])
        o.set_assignments([np.array(z, dtype=np.int32) for z in case["frozen_counts_z"]])
        ref = case["counts_and_histograms"]
        for m in range(M):
            nwk, nk = o.get_counts(m)
            assert np.array_equal(nwk, np.array(ref["typeTopicCounts"][m])) and np.array_equal(nk, np.array(ref["tokensPerTopic"][m]))
            want = np.array(ref["topicDocCounts"][m])
            got = o.get_hist(m)
            w = min(want.shape[1], got.shape[1])
            assert np.array_equal(got[:, 1:w], want[:, 1:w]) and not want[:, w:].any() and not got[:, w:].any(), (case["name"], m)
            b, bs = optim.optimize_beta(nwk, nk, Vs[m], case["beta"][m], case["betaSum"][m])
            assert b == pytest.approx(case["optimize_beta"]["beta"][m], rel=1e-12), (case["name"], m)
            assert bs == pytest.approx(case["optimize_beta"]["betaSum"][m], rel=1e-12), (case["name"], m)


def _trajectory():
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_trajectory.json")
    g = json.load(open(path))
    views = [(np.array(v["off"], dtype=np.int64), np.array(v["word"], dtype=np.int32)) for v in g["views"]]
    return g, g["K"], g["V"], views


def test_oracle_reproduces_reference_trajectory(oracle_mod):
    """30 sweeps (~360 K token draws) of the reference's own sampler + updater bytecode over a 400-document two-view corpus with
    the burn-in ramp of p_a (tests/golden/make_reference_trajectory.py): the oracle, fed the same uniforms, ends with the SAME
    assignments and the same log-likelihood at every checkpoint."""
    O = oracle_mod
    g, K, Vs, views = _trajectory()
    M = len(Vs)
    present = [np.ones(len(views[0][0]) - 1, dtype=np.uint8)] + [((v[0][1:] - v[0][:-1]) > 0).astype(np.uint8) for v in views[1:]]
    o = O.Oracle(K, Vs, views, seed=g["seed"], present=present)
    o.set_assignments([np.array(z, dtype=np.int32) for z in g["z0"]])
    o.rebuild_trees()
    marks = {it: ll for it, ll in g["loglik"]}
    assert np.allclose(o.loglik(True), marks[0], rtol=1e-12)
    for it in range(1, max(marks) + 1):
        o.set_hyper(p_a=np.full((M, M), min(it / 100.0 + 0.3, 1.1)))
        o.sweep(it, O.F_STALE_TREES | O.F_Q1_COMPAT)
        if it in marks:
            assert np.allclose(o.loglik(True), marks[it], rtol=1e-12), it
    for m in range(M):
        assert np.array_equal(o.get_assignments(m), np.array(g["z_final"][m], dtype=np.int32))


def test_flag_rule_equals_reference_dense_index(oracle_mod):
    """The engine's Q1 mode does not keep the reference's sorted list S; it keeps one flag per topic: "left the index (no view held
    it any more) or was gained while absent" -- and takes S = {held and not flagged}.  Inside the reference-faithful oracle (which
    reproduces the reference's sampler bytecode token for token) that rule is run next to the real list at every token of
    several sweeps: they must never disagree."""
    from helpers import random_corpus
    O = oracle_mod
    for K, Vs, means, D in [(12, [30, 10, 8], [7, 3, 2], 200), (37, [120, 30], [25, 5], 150), (50, [300], [6], 300)]:
        views = random_corpus(K, D, K, Vs, means)
        o = O.Oracle(K, Vs, views, seed=3)
        o.init_assignments(); o.rebuild_trees()
        for it in range(1, 7):
            if len(Vs) > 1:
                o.set_hyper(p_a=np.full((len(Vs), len(Vs)), 0.8))
            o.sweep(it, O.F_STALE_TREES | O.F_Q1_COMPAT | O.F_CHECK_RULE)
        assert o.rule_violations() == 0


def test_mallet_next_beta_law_matches_mallet_bytecode(oracle_mod):
    """The per-document view-coupling draw (W:333) in the reference is MALLET's Randoms.nextBeta.  Its law -- incl. quirk Q5: for
    alpha > 1, beta = 1 the first normal proposal is always accepted, a truncated N(1, 0.25/(alpha-1)) instead of Beta(alpha, 1) --
    as drawn by the MALLET jar's own bytecode (tests/golden/make_reference_beta_vectors.py) vs the oracle's restatement
    orc_next_beta_mallet (flag ORC_F_BETA_MALLET): two-sample Kolmogorov-Smirnov at alpha = 0.001."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_beta_vectors.json")))
    assert len(g["cases"]) >= 8
    for k, case in enumerate(g["cases"]):
        ref = np.sort(np.array(case["samples"]))
        got = np.sort(np.array([oracle_mod.next_beta_mallet(1000003 * k + s, case["a"], case["b"]) for s in range(6000)]))
        grid = np.concatenate([ref, got])
        d = np.abs(np.searchsorted(ref, grid, side="right") / len(ref) - np.searchsorted(got, grid, side="right") / len(got)).max()
        assert d < 1.95 * np.sqrt((len(ref) + len(got)) / (len(ref) * len(got))), (case["a"], case["b"], d)
        if case["a"] > 1 and case["b"] == 1:
            assert case["uniforms_per_draw"] == 1.0                    # Q5: never a second round of the rejection loop
            assert ref.mean() < case["a"] / (case["a"] + 1) - 0.01     # ... and not the Beta(a, 1) law


# ---- the inference path (FastQMVWVTopicInferencer) pinned to the executed jar ---------------------------------
def _inference_cases():
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_inference_vectors.json")))
    for case in gold["cases"]:
        views = [(np.array(v["off"], dtype=np.int64), np.array(v["word"], dtype=np.int32)) for v in case["views"]]
        yield case, case["K"], case["V"], views


def _inference_oracle(O, case, K, Vs, views):
    o = O.Oracle(K, Vs, views, seed=case["seed"])
    o.set_hyper(alpha=np.array(case["alpha"]), alphaSum=np.array(case["alphaSum"]), beta=np.array(case["beta"]),
                betaSum=np.array(case["betaSum"]), gamma=np.array(case["gamma"]), p_a=np.array(case["p_a"]), p_b=np.array(case["p_b"]), inactive=[])
    for m in range(len(Vs)):
        o.set_counts(m, np.array(case["n_wk"][m], dtype=np.int32), np.array(case["n_k"][m], dtype=np.int32))
    return o


def test_inference_path_matches_reference_bytecode(oracle_mod):
    """What FastQMVWVTopicInferencer does (I:114-330) as executed from the shipped jar (tests/golden/
    make_reference_inference_vectors.py: `new FTree(phi)` per word, FTree.sample for every in-vocabulary token's first topic,
    sampleTopicsForOneDoc with nut = 0 and those gamma*alpha-less trees) vs the oracle's inference mode: initial assignments
    (incl. out-of-vocabulary tokens keeping topic 0) and the assignments after every frozen sweep TOKEN FOR TOKEN; the trained
    tables never move."""
    O = oracle_mod
    n_tok = n_oov = 0
    for case, K, Vs, views in _inference_cases():
        M = len(Vs)
        o = _inference_oracle(O, case, K, Vs, views)
        o.init_from_phi()
        for m in range(M):
            want = np.array(case["z_init"][m], dtype=np.int32)
            assert np.array_equal(o.get_assignments(m), want), (case["name"], "init", m)
            oov = views[m][1] >= Vs[m]
            assert np.all(want[oov] == 0)
            n_oov += int(oov.sum())
        for it, want in enumerate(case["z_after"], start=1):
            o.sweep(it, O.F_FROZEN | O.F_BARE_TREES | O.F_Q1_COMPAT)
            for m in range(M):
                got, w = o.get_assignments(m), np.array(want[m], dtype=np.int32)
                assert np.array_equal(got, w), (case["name"], "sweep", it, "view", m, "first mismatch at", int(np.argmax(got != w)))
                n_tok += len(w)
        for m in range(M):
            nwk, nk = o.get_counts(m)
            assert np.array_equal(nwk, np.array(case["n_wk"][m])) and np.array_equal(nk, np.array(case["n_k"][m]))
    assert n_tok > 1500 and n_oov > 10


def test_inference_conditionals_match_reference_bytecode(oracle_mod):
    """Per-token conditionals of the inference path on the trained (frozen) tables as the jar's sampler computed them with the
    inferencer's trees (document masses + phi leaves, no gamma*alpha: quirk Q13), vs the oracle's three-bucket masses and its
    engine-form dense distribution in bare-tree mode: 1e-9 relative.  Q1-affected tokens included."""
    O = oracle_mod
    n = n_q1 = 0
    for case, K, Vs, views in _inference_cases():
        if not case["conditionals"]:
            continue
        M = len(Vs)
        o = _inference_oracle(O, case, K, Vs, views)
        base = [np.array(z, dtype=np.int32) for z in case["z_after"][-1]]
        counts = [(np.array(case["n_wk"][m], dtype=np.int32), np.array(case["n_k"][m], dtype=np.int32)) for m in range(M)]
        for rec in case["conditionals"][::2]:
            zs = [z.copy() for z in base]
            for m, zd in enumerate(rec["z_doc"]):
                if zd is not None:
                    b = int(views[m][0][rec["doc"]])
                    zs[m][b:b + len(zd)] = zd
            o.set_assignments(zs)                       # (rebuilds counts from z: put the trained tables back)
            for m in range(M):
                o.set_counts(m, *counts[m])
            p = np.eye(M); p[rec["view"]] = rec["p_row"]
            for engine_form in (False, True):
                got = o.cond_probs(rec["view"], rec["doc"], rec["pos"], p=p, engine_form=engine_form, not_in_S=rec.get("not_in_S"),
                                   flags=O.F_BARE_TREES)
                assert np.allclose(got[:K], rec["probs"], rtol=1e-9, atol=1e-15), (case["name"], rec["doc"], rec["view"], rec["pos"], engine_form)
                assert got[K] == 0.0 and rec["new_share"] == 0.0          # the inferencer's inactive set is empty (I:243)
            n += 1
            n_q1 += bool(rec.get("not_in_S"))
    assert n > 250
