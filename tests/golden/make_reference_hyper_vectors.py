"""Samples drawn BY THE REFERENCE'S OWN SAMPLER BYTECODE (org.knowceans.util.RandomSamplers / Samplers inside
/root/reference/output/MVTopicModel-1.0-SNAPSHOT.jar), executed without a JVM by tools/jvm_mini.py.  The only thing replaced is
the source of uniforms: java.util.Random.nextDouble and Cokus.randDouble are served from a seeded numpy stream.

  RandomSamplers.randGamma(shape, scale)   KR:294-360   the calls of optimizeGamma (M:2394, M:2410, M:2424)
  RandomSamplers.randBeta(a, b)            KR:267-271   (M:2388, M:2404, M:2420)
  RandomSamplers.randBernoulli(p)          KR:789-795   (M:2393, M:2409, M:2418)
  Randoms.nextGamma(alpha, 1)  (MALLET)    the Gamma draws of sampleDirichlet (M:2616), from output/lib/mallet-2.0.8.jar
  Samplers.randAntoniak(alpha, n)          KS:1089-1110 (M:2471, M:2500): the FIRST call for a given n on a fresh class inverts the
                                           exact Stirling-number law; every later call reads a cache row the earlier calls
                                           multiplied and prefix-summed in place (quirk Q7) -- both are recorded.

Output: tests/golden/reference_hyper_vectors.json; tests/test_optim_host.py compares the product's host samplers with it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import jvm_mini  # noqa: E402

REF = "/root/reference/output"
RS, KS = "org/knowceans/util/RandomSamplers", "org/knowceans/util/Samplers"


def fresh_vm(uniform):
    vm = jvm_mini.MiniJVM([os.path.join(REF, "MVTopicModel-1.0-SNAPSHOT.jar")])
    vm.shims["java/util/Random.nextDouble:()D"] = lambda loc, recv, args, pc: uniform()
    vm.shims["org/knowceans/util/Cokus.randDouble:()D"] = lambda loc, recv, args, pc: uniform()
    return vm


def r7(x):
    return float(f"{x:.7g}")


def main():
    rng = np.random.default_rng(20261018)
    vm = fresh_vm(lambda: float(rng.random()))
    samp = vm.new(RS, "(Ljava/util/Random;)V", [jvm_mini.JObject("java/util/Random")])
    N = 1500
    out = {"source": "drawn by the reference's sampler bytecode under tools/jvm_mini.py; uniforms from numpy default_rng(20261018)",
           "randGamma": [], "randBeta": [], "randBernoulli": [], "randAntoniak_first_call": [], "randAntoniak_repeated": []}
    for shape, scale in [(0.3, 2.0), (1.0, 1.0), (3.0, 0.5), (40.0, 0.1), (250.0, 1.0 / 37.5)]:
        out["randGamma"].append({"shape": shape, "scale": scale,
                                 "samples": [r7(vm.call(RS, "randGamma", "(DD)D", [samp, shape, scale])) for _ in range(N)]})
    for a, b in [(1.0, 1.0), (2.0, 5.0), (0.3, 1.0), (11.0, 250.0), (1.5, 40.0)]:
        out["randBeta"].append({"a": a, "b": b, "samples": [r7(vm.call(RS, "randBeta", "(DD)D", [samp, a, b])) for _ in range(N)]})
    for p in (0.0, 0.05, 0.3, 0.9, 1.0):
        out["randBernoulli"].append({"p": p, "n": 4000, "ones": int(sum(vm.call(RS, "randBernoulli", "(D)I", [samp, p]) for _ in range(4000)))})
    # Antoniak: one call per fresh class (the cache is static), at chosen uniforms -> a deterministic statement of the law it inverts
    us = [0.0, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 0.999999]
    for alpha, n in [(0.1, 2), (0.1, 7), (1.0, 30), (3.7, 12), (0.02, 200), (25.0, 60)]:
        rec = {"alpha": alpha, "n": n, "draws": []}
        for u in us:
            v = fresh_vm(lambda u=u: u)
            rec["draws"].append([u, v.call(KS, "randAntoniak", "(DI)I", [alpha, n])])
        out["randAntoniak_first_call"].append(rec)
    # ... and the same arguments over and over on one class: Q7 at work
    for alpha, n in [(3.7, 12), (0.5, 40)]:
        v = fresh_vm(lambda: float(rng.random()))
        xs = [v.call(KS, "randAntoniak", "(DI)I", [alpha, n]) for _ in range(400)]
        out["randAntoniak_repeated"].append({"alpha": alpha, "n": n, "draws": xs,
                                             "exact_mean": float(sum(alpha / (alpha + i) for i in range(n)))})
    # MALLET's Randoms.nextGamma(alpha, 1) -- the Gamma draws of sampleDirichlet (M:2616) -- from the MALLET jar
    mv = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar")])
    MR = "cc/mallet/util/Randoms"
    mv.shims[MR + ".nextUniform:()D"] = lambda loc, r, a, pc: float(rng.random())
    mv.shims[MR + ".nextGaussian:()D"] = lambda loc, r, a, pc: float(rng.standard_normal())
    mobj = jvm_mini.JObject(MR)
    out["mallet_nextGamma"] = [{"shape": a, "samples": [r7(mv.call(MR, "nextGamma", "(DD)D", [mobj, a, 1.0])) for _ in range(N)]}
                               for a in (0.05, 0.5, 1.0, 2.5, 40.0, 900.0)]
    json.dump(out, open(os.path.join(HERE, "reference_hyper_vectors.json"), "w"))
    print("reference_hyper_vectors.json written;", {k: len(v) for k, v in out.items() if isinstance(v, list)})
    for r in out["randAntoniak_repeated"]:
        print("  Q7: randAntoniak(%g, %d) repeated: mean %.3f, exact law %.3f" % (r["alpha"], r["n"], np.mean(r["draws"]), r["exact_mean"]))


if __name__ == "__main__":
    main()
