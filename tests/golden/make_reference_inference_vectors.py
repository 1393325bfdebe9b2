"""Golden vectors of the INFERENCE path (FastQMVWVTopicInferencer I:114-330) produced by executing the
reference's own binary.

The inferencer class itself is absent from the shipped jar (output/MVTopicModel-1.0-SNAPSHOT.jar predates it), but everything it
DOES is code that is in the jar:
  * I:561-576 initInferencer builds one `new FTree(temp)` per word with temp[t] = (n_wk[w][t] + beta) / (n_k[t] + betaSum) -- bare
    phi, no gamma*alpha (quirk Q13).  Here: temp is computed by these five lines of Python, FTree.<init> is EXECUTED from the jar;
  * I:186-203 draws every in-vocabulary token's first topic with `trees[m][type].sample(u)` (out-of-vocabulary tokens keep the
    0 of `new int[tokens.size()]`).  Here: FTree.sample EXECUTED from the jar, u = the draw the oracle / engine use for
    (position, document, view, PURPOSE_INIT);
  * I:211-256 runs FastQMVWVWorkerRunnable with nst = 1, nut = 0, queues = null, an EMPTY inActiveTopicIndex and p_a = 0.2,
    p_b = 1: sampleTopicsForOneDoc EXECUTED from the jar (tests/golden/make_reference_sampler_vectors.py::RefSampler) with the
    trained tables installed, those trees, and deltas dropped (W:587 `nut > 0` is false, so nothing reaches a queue).
Recorded: the initial assignments, the assignments after each of the sweeps, and -- one more sweep, frozen -- the per-token
conditionals read out of the sampler's frame (as reference_conditionals does for the trainer).

Output: tests/golden/reference_inference_vectors.json.  Needs /root/reference (build container only).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import make_reference_sampler_vectors as R  # noqa: E402
from jvm_mini import JObject  # noqa: E402
from oracle import oracle as O  # noqa: E402

PURPOSE_INIT = 1


def make_case(name, K, Vs, means_train, means_new, D_train, D_new, seed, train_sweeps, sweeps, rng_seed, nonuniform_alpha=False):
    from helpers import random_corpus
    M = len(Vs)
    # a trained state: any reachable count tables will do; the sequential oracle makes them
    train = random_corpus(rng_seed, D_train, K, Vs, means_train, empty_frac=0.1)
    rng = np.random.default_rng(rng_seed + 100)
    alpha = np.full((M, K + 1), 0.1)
    if nonuniform_alpha:
        alpha = rng.uniform(0.02, 0.3, size=(M, K + 1))
    alphaSum = alpha[:, :K].sum(1) if nonuniform_alpha else np.full(M, 0.1 * K)
    gamma = rng.uniform(0.6, 1.6, size=M) if nonuniform_alpha else np.ones(M)
    beta = np.full(M, 0.01)
    t = O.Oracle(K, Vs, train, seed=seed)
    t.set_hyper(alpha=alpha, alphaSum=alphaSum, gamma=gamma)
    t.init_assignments(); t.rebuild_trees()
    for it in range(1, train_sweeps + 1):
        t.sweep(it, O.F_STALE_TREES)
    counts = [t.get_counts(m) for m in range(M)]
    # new documents.  Out-of-vocabulary word ids (type >= numTypes[m]) are skipped by I:188 and by W:427-428 of the SOURCE, but the
    # shipped jar's build of the sampler predates W:427-428 (its bytecode indexes typeTopicCounts[m][type] unguarded: offset 792),
    # so only the initialisation-only case (sweeps = 0) carries OOV ids
    new = random_corpus(rng_seed + 1, D_new, K, Vs, means_new, empty_frac=0.1, oov=(sweeps == 0))
    z_zero = [[0] * len(v[1]) for v in new]                     # `new int[tokens.size()]`
    pa, pb = np.full((M, M), 0.2), np.ones((M, M))              # I:227-230
    # RefSampler's constructor runs the TRAINER's buildInitialTypeTopicCounts (M:600-652), which indexes typeTopicCounts[m][type]
    # and would throw on an out-of-vocabulary id; the inferencer never calls it.  So the object is built over in-vocabulary
    # stand-ins, and the documents' real word lists are put back (in place) before anything of the inference path runs.
    clean = [(off, np.where(w < V, w, 0).astype(np.int32)) for (off, w), V in zip(new, Vs)]
    ref = R.RefSampler(K, Vs, clean, z_zero, seed, alpha, alphaSum, beta, gamma, pa, pb, inactive=())
    ref.views = new
    for d, ent in enumerate(ref.docs):
        for m, ta in enumerate(ent.fields["Assignments"]):
            if ta is not None:
                b, e = int(new[m][0][d]), int(new[m][0][d + 1])
                ta.fields["instance"][1][:] = [int(x) for x in new[m][1][b:e]]
    vm = ref.vm
    # the trained tables replace what buildInitialTypeTopicCounts made of the all-zero assignments (in place: the worker object
    # holds these very lists)
    for m in range(M):
        for w in range(Vs[m]):
            ref.nwk[m][w][:] = [int(x) for x in counts[m][0][w]]
        ref.nk[m][:] = [int(x) for x in counts[m][1]]
    # I:561-576 -- bare-phi trees through the jar's FTree constructor
    for m in range(M):
        for w in range(Vs[m]):
            temp = [(ref.nwk[m][w][k] + ref.beta[m]) / (ref.nk[m][k] + ref.betaSum[m]) for k in range(K)]
            tree = JObject(R.FT)
            vm.call(R.FT, "<init>", "([D)V", [tree, temp])
            ref.trees[m][w] = tree
    # I:186-203 -- first topics by FTree.sample (jar), entity by entity, view by view, position by position
    for d, ent in enumerate(ref.docs):
        for m, ta in enumerate(ent.fields["Assignments"]):
            if ta is None:
                continue
            words, topics = ta.fields["instance"][1], ta.fields["topicSequence"][3]
            for pos, w in enumerate(words):
                if not (w < Vs[m]):
                    continue
                x = O.philox([pos, d, 0, (m << 8) | PURPOSE_INIT], [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
                topics[pos] = vm.call(R.FT, "sample", "(D)I", [ref.trees[m][w], R.u24(x[0])])
    z_init = [[0] * len(v[1]) for v in new]
    for d, ent in enumerate(ref.docs):
        for m, ta in enumerate(ent.fields["Assignments"]):
            if ta is not None:
                _, _, b, zs = ta.fields["topicSequence"]
                z_init[m][b:b + len(zs)] = zs
    # nut = 0: nothing is enqueued (W:587); in this build of the worker the enqueue is unconditional, so the queue drops it
    vm.shims["java/util/Queue.add:(Ljava/lang/Object;)Z"] = lambda loc, r, a, pc: 1
    out = {"name": name, "K": K, "V": Vs, "seed": seed, "views": [{"off": v[0].tolist(), "word": v[1].tolist()} for v in new],
           "alpha": alpha.tolist(), "alphaSum": np.asarray(alphaSum).tolist(), "beta": beta.tolist(), "betaSum": ref.betaSum,
           "gamma": np.asarray(gamma).tolist(), "p_a": pa.tolist(), "p_b": pb.tolist(),
           "n_wk": [c[0].tolist() for c in counts], "n_k": [c[1].tolist() for c in counts], "z_init": z_init, "z_after": []}
    for it in range(1, sweeps + 1):
        out["z_after"].append(ref.sweep(it))
    for m in range(M):      # the tables did not move
        assert all(ref.nwk[m][w] == [int(x) for x in counts[m][0][w]] for w in range(Vs[m])) and ref.nk[m] == [int(x) for x in counts[m][1]]
    out["conditionals"] = R.reference_conditionals(ref, sweeps + 1, max_tokens=400, rebuild=False) if sweeps else []
    out["counters"] = dict(ref.counters)
    print(name, "tokens", [len(z) for z in z_init], "sweeps", sweeps, "conditionals", len(out["conditionals"]), ref.counters,
          "bytecode steps", vm.steps, flush=True)
    return out


def main():
    cases = [
        make_case("infer_single_view", K=8, Vs=[30], means_train=[7], means_new=[6], D_train=60, D_new=30, seed=21, train_sweeps=5,
                  sweeps=3, rng_seed=11),
        make_case("infer_two_views", K=10, Vs=[40, 12], means_train=[8, 3], means_new=[7, 3], D_train=60, D_new=25, seed=22,
                  train_sweeps=5, sweeps=3, rng_seed=12, nonuniform_alpha=True),
        make_case("infer_three_views_k37", K=37, Vs=[90, 20, 12], means_train=[14, 4, 3], means_new=[12, 4, 2], D_train=50, D_new=16,
                  seed=23, train_sweeps=4, sweeps=2, rng_seed=13, nonuniform_alpha=True),
        make_case("infer_init_only_with_oov", K=21, Vs=[60, 15], means_train=[10, 4], means_new=[30, 12], D_train=50, D_new=40, seed=24,
                  train_sweeps=4, sweeps=0, rng_seed=14),
    ]
    json.dump({"source": "FTree.<init> / FTree.sample / FastQMVWVWorkerRunnable.sampleTopicsForOneDoc from output/MVTopicModel-1.0-SNAPSHOT.jar, "
                         "driven as FastQMVWVTopicInferencer I:114-330 drives them (bare-phi trees, nut = 0), executed by tools/jvm_mini.py",
               "cases": cases}, open(os.path.join(HERE, "reference_inference_vectors.json"), "w"))


if __name__ == "__main__":
    main()
