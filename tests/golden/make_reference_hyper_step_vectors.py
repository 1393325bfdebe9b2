"""optimizeDP (M:2440-2591) and optimizeGamma (M:2369-2438) EXECUTED from the reference's jar
(/root/reference/output/MVTopicModel-1.0-SNAPSHOT.jar) by tools/jvm_mini.py, with the sampler calls SCRIPTED: every call of

    Samplers.randAntoniak(alpha, n)            RandomSamplers.randBeta(a, b)      RandomSamplers.randBernoulli(p)
    RandomSamplers.randGamma(shape, scale)     Randoms.nextGamma(shape, 1)  (inside sampleDirichlet, M:2593-2632)

is answered by a value drawn here from the same law (numpy) and written down, in call order, together with the arguments the
bytecode passed.  The product's host code (mvtm_test_hyper_core in libmvtm.so: the very functions mvtm_optimize_hyper runs) is then
fed the same values: it must ask for the same draws with the same arguments and end with the same alpha / alphaSum / tablesCnt /
rootTablesCnt / gammaRoot / gammaView / gamma / inactive topics (tests/test_optim_host.py).  Degenerate calls the engine skips
(Bernoulli(0) = 0 and Beta(a, 0) = 1 for the zero-length bin of docLengthCounts) are answered without an entry.

Output: tests/golden/reference_hyper_step_vectors.json
"""
import copy
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import jvm_mini  # noqa: E402
from jvm_mini import JObject  # noqa: E402

REF = "/root/reference/output"
MC = "org/madgik/MVTopicModel/FastQMVWVParallelTopicModel"


def run_case(name, M, K, hist, lencnt, alpha, gamma, gammaView, gammaRoot, seed):
    rng = np.random.default_rng(seed)
    vm = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar"), os.path.join(REF, "MVTopicModel-1.0-SNAPSHOT.jar")])
    script = []                                          # [kind, a, b, value]: kind 1 Gamma(a,1) 2 Beta 3 Bernoulli 4 Antoniak
    inactive = []
    sh = vm.shims

    def antoniak(loc, r, a, pc):
        alpha_, n = a
        tables = int(sum(rng.random() < alpha_ / (alpha_ + i) for i in range(n)))
        script.append([4, alpha_, n, max(tables, 1)])
        return max(tables, 1)

    def beta(loc, r, a, pc):
        if a[1] == 0:
            return 1.0                                   # KR:267-271: randGamma(0) = 0 -> x / (x + 0)
        v = float(rng.beta(a[0], a[1]))
        v = min(max(v, 1e-300), 1.0)
        script.append([2, a[0], a[1], v])
        return v

    def bernoulli(loc, r, a, pc):
        if a[0] <= 0:
            return 0
        v = int(rng.random() < a[0])
        script.append([3, a[0], 0, v])
        return v

    def gamma2(loc, r, a, pc):                           # randGamma(shape, scale) = randGamma(shape) * scale  (KR:358-360)
        unit = float(rng.gamma(a[0]))
        script.append([1, a[0], a[1], unit])
        return unit * a[1]

    sh["org/knowceans/util/Samplers.randAntoniak:(DI)I"] = antoniak
    sh["org/knowceans/util/RandomSamplers.randBeta:(DD)D"] = beta
    sh["org/knowceans/util/RandomSamplers.randBernoulli:(D)I"] = bernoulli
    sh["org/knowceans/util/RandomSamplers.randGamma:(DD)D"] = gamma2
    sh["cc/mallet/util/Randoms.nextGamma:(DD)D"] = gamma2
    unbox = lambda v: v.fields["value"] if isinstance(v, JObject) else v

    def list_remove(loc, r, a, pc):
        v = unbox(a[0])
        if v in inactive:
            inactive.remove(v); return 1
        return 0
    sh["java/util/List.add:(Ljava/lang/Object;)Z"] = lambda loc, r, a, pc: (inactive.append(unbox(a[0])), 1)[1]
    sh["java/util/List.remove:(Ljava/lang/Object;)Z"] = list_remove
    sh["java/util/List.isEmpty:()Z"] = lambda loc, r, a, pc: int(len(inactive) == 0)
    sh["java/util/List.size:()I"] = lambda loc, r, a, pc: len(inactive)
    sh["java/util/List.get:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: inactive[a[0]]
    sh["java/lang/Integer.valueOf:(I)Ljava/lang/Integer;"] = lambda loc, r, a, pc: a[0]

    def integer_init(loc, r, a, pc):
        r.fields["value"] = a[0]
    sh["java/lang/Integer.<init>:(I)V"] = integer_init
    sh["java/lang/Byte.valueOf:(B)Ljava/lang/Byte;"] = lambda loc, r, a, pc: a[0]
    sh["java/lang/Byte.byteValue:()B"] = lambda loc, r, a, pc: r
    for d in ("(Ljava/lang/String;)Ljava/lang/StringBuilder;", "(D)Ljava/lang/StringBuilder;", "(I)Ljava/lang/StringBuilder;",
              "(Ljava/lang/Object;)Ljava/lang/StringBuilder;"):
        sh["java/lang/StringBuilder.append:" + d] = lambda loc, r, a, pc: r
    sh["java/lang/StringBuilder.toString:()Ljava/lang/String;"] = lambda loc, r, a, pc: ""
    sh["org/apache/log4j/Logger.info:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: None
    sh["java/text/NumberFormat.format:(D)Ljava/lang/String;"] = lambda loc, r, a, pc: ""
    sh["java/text/NumberFormat.format:(Ljava/lang/Object;)Ljava/lang/String;"] = lambda loc, r, a, pc: ""
    vm.statics[(MC, "logger")] = JObject("logger")

    model = JObject(MC)
    model.fields.update(dict(numModalities=M, numTopics=K, alpha=copy.deepcopy(alpha), alphaSum=[float(sum(a)) for a in alpha],
                             gamma=list(gamma), gammaView=list(gammaView), gammaRoot=float(gammaRoot), rootTablesCnt=0.0,
                             tablesCnt=[0.0] * M, topicDocCounts=copy.deepcopy(hist), docLengthCounts=copy.deepcopy(lencnt),
                             inActiveTopicIndex=("inactive",), samp=JObject("org/knowceans/util/RandomSamplers"),
                             random=("randoms",), formatter=("nf",)))
    rec = {"name": name, "M": M, "K": K, "topicDocCounts": hist, "docLengthCounts": lencnt,
           "in": {"alpha": alpha, "gamma": list(gamma), "gammaView": list(gammaView), "gammaRoot": float(gammaRoot)}}
    vm.strict_fields = True
    vm.call(MC, "optimizeDP", "()V", [model])
    n_dp = len(script)
    f = model.fields
    rec["after_optimizeDP"] = {"alpha": copy.deepcopy(f["alpha"]), "alphaSum": list(f["alphaSum"]), "tablesCnt": list(f["tablesCnt"]),
                               "rootTablesCnt": f["rootTablesCnt"], "inactive": sorted(inactive), "draws": n_dp}
    vm.call(MC, "optimizeGamma", "()V", [model])
    rec["after_optimizeGamma"] = {"gammaRoot": f["gammaRoot"], "gammaView": list(f["gammaView"]), "gamma": list(f["gamma"]),
                                  "draws": len(script) - n_dp}
    rec["script"] = script
    print(f"  {name}: {n_dp} + {len(script) - n_dp} scripted draws, bytecode steps {vm.steps}, gamma {f['gamma']}, gammaRoot {f['gammaRoot']:.4f}")
    return rec


def synth(rng, M, K, D, mean_len, dead=()):
    """topicDocCounts / docLengthCounts of a synthetic assignment: D documents, Poisson lengths, topics from a skewed law"""
    hist, lencnt = [], []
    for m in range(M):
        lens = rng.poisson(mean_len[m], D)
        if m > 0:
            lens[rng.random(D) < 0.3] = 0                  # the side view is missing in some documents
        w = rng.dirichlet(np.full(K, 0.3))
        w[list(dead)] = 0.0
        w /= w.sum()
        stride = int(lens.max()) + 1
        h = np.zeros((K, stride), dtype=np.int64)
        for L in lens:
            if L:
                theta = rng.dirichlet(w * 3 + 1e-9)
                c = np.bincount(rng.choice(K, size=L, p=theta), minlength=K)
                for t in np.nonzero(c)[0]:
                    h[t, c[t]] += 1
        hist.append(h.tolist())
        lc = np.bincount(lens[lens > 0] if m > 0 else lens, minlength=stride)   # absent views are not counted (M:626 counts present docs)
        lencnt.append(lc.tolist())
    return hist, lencnt


def main():
    rng = np.random.default_rng(20261018)
    out = {"source": "optimizeDP / optimizeGamma executed from the reference's jar by tools/jvm_mini.py with scripted sampler calls",
           "kinds": {"1": "Gamma(a,1) (b = the scale the caller applies)", "2": "Beta(a,b)", "3": "Bernoulli(a)", "4": "Antoniak(a, n=b)"},
           "cases": []}
    K = 12
    hist, lencnt = synth(rng, 1, K, 60, [9.0])
    alpha = [(np.full(K + 1, 1.0 / (K + 1))).tolist()]
    out["cases"].append(run_case("one_view", 1, K, hist, lencnt, alpha, [1.0], [1.0], 10.0, 1))
    K = 9
    hist, lencnt = synth(rng, 2, K, 80, [14.0, 3.0], dead=(4, 7))
    alpha = [rng.dirichlet(np.full(K + 1, 2.0)).tolist() for _ in range(2)]
    out["cases"].append(run_case("two_views_dead_topics", 2, K, hist, lencnt, alpha, [1.7, 0.6], [2.5, 0.8], 4.2, 2))
    K = 6
    hist, lencnt = synth(rng, 3, K, 40, [6.0, 2.0, 1.2])
    lencnt[0][0] = 3                                         # zero-length documents present in view 0: the j = 0 bin of M:2415-2422
    alpha = [rng.dirichlet(np.full(K + 1, 0.7)).tolist() for _ in range(3)]
    out["cases"].append(run_case("three_views_zero_length_bin", 3, K, hist, lencnt, alpha, [0.9, 2.4, 1.1], [1.0, 1.0, 3.0], 10.0, 3))
    json.dump(out, open(os.path.join(HERE, "reference_hyper_step_vectors.json"), "w"))
    print("reference_hyper_step_vectors.json:", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
