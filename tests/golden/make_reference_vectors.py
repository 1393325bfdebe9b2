"""Golden vectors produced BY THE REFERENCE'S OWN BINARIES, executed here without a JVM through tools/jvm_mini.py
(a bytecode interpreter for numeric leaf methods).  Run in the build container (needs /root/reference):

  /root/reference/output/MVTopicModel-1.0-SNAPSHOT.jar   org.madgik.utils.FTree (constructor, sample, update)
                                                         FastQMVWVWorkerRunnable.lower_bound  (this prebuilt jar is an older build
                                                         of the class: lower_bound takes int[]; the algorithm is W:257-277)
  /root/reference/output/lib/mallet-2.0.8.jar            cc.mallet.types.Dirichlet.logGammaStirling / digamma /
                                                         learnSymmetricConcentration   (pom.xml:49-53 pins this version)

Output: tests/golden/reference_vectors.json.  tests/test_oracle.py and tests/test_optim_host.py compare the C oracle, the numpy
restatement and the product's host code against it.
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


def main():
    vm = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar"), os.path.join(REF, "MVTopicModel-1.0-SNAPSHOT.jar")])
    rng = np.random.default_rng(20261018)
    out = {"source": "executed from the reference's jars by tools/jvm_mini.py", "ftree": [], "lower_bound": [], "logGammaStirling": [],
           "digamma": [], "learnSymmetricConcentration": []}
    FT = "org/madgik/utils/FTree"
    us = [0.0, 1e-12, 0.0999, 0.1, 0.25, 0.3, 0.4, 0.5, 0.6, 0.75, 0.9, 0.999999, 1.0] + rng.random(20).tolist()
    for K in (4, 5, 7, 8, 37, 50, 500):
        w = [1.0, 2.0, 3.0, 4.0] if K == 4 else (rng.gamma(0.3, 1.0, K) + (rng.random(K) < 0.2) * 0.0).tolist()
        if K == 7:
            w[2] = 0.0; w[5] = 0.0                                  # zero-weight leaves
        t = vm.new(FT, "([D)V", [list(w)])
        rec = {"K": K, "weights": list(w), "tree": list(t.fields["tree"]),
               "samples": [[u, vm.call(FT, "sample", "(D)I", [t, u])] for u in us], "updates": []}
        for _ in range(6):
            topic, v = int(rng.integers(0, K)), float(rng.gamma(0.5, 2.0))
            vm.call(FT, "update", "(ID)V", [t, topic, v])
            rec["updates"].append({"topic": topic, "value": v, "tree": list(t.fields["tree"]),
                                   "samples": [[u, vm.call(FT, "sample", "(D)I", [t, u])] for u in us[:12]]})
        out["ftree"].append(rec)
    W = "org/madgik/MVTopicModel/FastQMVWVWorkerRunnable"
    for arr in ([1, 3, 6, 10], [5], [2, 2, 2, 9], sorted(rng.integers(0, 50, 17).tolist())):
        for key in [0.5, 1.0, 1.01, 2.0, 6.0, 9.0, 10.0, 10.5, 49.0, 60.0] + rng.uniform(0, 50, 5).tolist():
            for n in {len(arr), max(1, len(arr) - 1)}:
                out["lower_bound"].append({"arr": arr, "key": key, "n": n, "result": vm.call(W, "lower_bound", "([IDI)I", [list(arr), key, n])})
    D = "cc/mallet/types/Dirichlet"
    for z in [1e-9, 1e-6, 0.0001, 0.01, 0.1, 0.5, 0.99, 1.0, 1.5, 2.0, 2.1, 3.0, 5.0, 9.4, 9.5, 9.6, 10.0, 100.5, 1234.5, 1e6] + rng.gamma(1.0, 5.0, 15).tolist():
        out["logGammaStirling"].append([z, vm.call(D, "logGammaStirling", "(D)D", [z])])
        out["digamma"].append([z, vm.call(D, "digamma", "(D)D", [z])])
    for case in range(6):
        ncount, nsize = int(rng.integers(3, 40)), int(rng.integers(5, 400))
        count_hist = [0] + rng.integers(0, 50, ncount).tolist()
        size_hist = rng.integers(0, 3, nsize + 1).tolist()
        size_hist[-1] = max(1, size_hist[-1])
        if case == 5:                                               # widely spaced non-empty lengths: the "gap > 20" branch
            size_hist = [0] * 300; size_hist[3] = 4; size_hist[40] = 2; size_hist[299] = 1
        V, start = int(rng.integers(50, 5000)), float(rng.uniform(0.5, 60.0))
        r = vm.call(D, "learnSymmetricConcentration", "([I[IID)D", [list(count_hist), list(size_hist), V, start])
        out["learnSymmetricConcentration"].append({"countHistogram": count_hist, "topicSizeHistogram": size_hist, "numDimensions": V,
                                                   "currentValue": start, "result": r})
    json.dump(out, open(os.path.join(HERE, "reference_vectors.json"), "w"))
    print("reference_vectors.json:", {k: len(v) for k, v in out.items() if isinstance(v, list)}, "bytecode steps", vm.steps)


if __name__ == "__main__":
    main()
