"""cc.mallet.util.Randoms.nextBeta(alpha, beta) -- the per-document view-coupling draw of the sampler (W:333) -- EXECUTED from
/root/reference/output/lib/mallet-2.0.8.jar by tools/jvm_mini.py, its nextUniform() / nextGaussian() served from a seeded numpy
stream.  1 500 draws per parameter pair.  This is quirk Q5 on record: for alpha > 1 and beta = 1 (the only beta the reference
passes) the acceptance test of the normal-proposal branch evaluates 0 * log(inf) = NaN, the comparison is false, and the first
proposal is returned: a normal N(1, 0.25/(alpha-1)) truncated to [0, 1], not Beta(alpha, 1).

Output: tests/golden/reference_beta_vectors.json; tests/test_reference_vectors.py compares the oracle's restatement
(orc_next_beta_mallet, flag ORC_F_BETA_MALLET) with it.
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
R = "cc/mallet/util/Randoms"


def main():
    rng = np.random.default_rng(20261018)
    vm = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar")])
    calls = {"u": 0, "g": 0}

    def uni(loc, r, a, pc):
        calls["u"] += 1
        return float(rng.random())

    def gau(loc, r, a, pc):
        calls["g"] += 1
        return float(rng.standard_normal())
    vm.shims[R + ".nextUniform:()D"] = uni
    vm.shims[R + ".nextGaussian:()D"] = gau
    obj = jvm_mini.JObject(R)
    out = {"source": "Randoms.nextBeta executed from mallet-2.0.8.jar by tools/jvm_mini.py; uniforms / normals from numpy default_rng(20261018)",
           "cases": []}
    for a, b in [(0.2, 1.0), (0.3, 1.0), (0.7, 1.0), (1.0, 1.0), (1.1, 1.0), (2.0, 1.0), (5.0, 1.0), (100.0, 1.0), (2.0, 3.0), (0.5, 0.5)]:
        calls["u"] = calls["g"] = 0
        xs = [vm.call(R, "nextBeta", "(DD)D", [obj, a, b]) for _ in range(1500)]
        out["cases"].append({"a": a, "b": b, "samples": [float(f"{x:.7g}") for x in xs], "uniforms_per_draw": calls["u"] / 1500,
                             "normals_per_draw": calls["g"] / 1500})
        print(f"  nextBeta({a}, {b}): mean {np.mean(xs):.4f} (Beta law: {a / (a + b):.4f}), {calls['u'] / 1500:.2f} uniforms and {calls['g'] / 1500:.2f} normals per draw")
    json.dump(out, open(os.path.join(HERE, "reference_beta_vectors.json"), "w"))


if __name__ == "__main__":
    main()
