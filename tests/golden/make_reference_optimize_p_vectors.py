"""optimizeP(boolean) EXECUTED from the reference's jar (/root/reference/output/MVTopicModel-1.0-SNAPSHOT.jar) by tools/jvm_mini.py.

The shipped jar is OLDER than the source for this method: it visits the views of a document in index order (view m against the
views i < m), where the source (M:2717-2741) orders them by descending length through a TreeMap.  On documents whose view lengths
DEcrease with the view index (text > keywords > ..., a missing view counts as 0) the two orders coincide, so on such corpora the
jar's output is what the source computes: the per-pair sums of pDistr_Mean (M:2706-2782), pMean, and the Beta parameters
p_a = min(-1/ln(pMean), 100), p_b = 1 (M:2785-2812).  Second difference: the jar divides the sum over documents by
totalDocsPerModality[m], the source by min(totalDocsPerModality[m], totalDocsPerModality[i]) (M:2793) -- equal when every document
has every view (cases marked "all_views_present": pMean and p_a comparable as they are); otherwise the test multiplies the jar's
pMean back by totalDocsPerModality[m] and compares the SUMS.  The TreeMap collision rule for EQUAL lengths (quirk Q11) is not
covered by these vectors; it rests on the source citation.

Output: tests/golden/reference_optimize_p.json; tests/test_optim_host.py compares oracle/optim.py (p_statistics, p_params), the
restatement the engine's k_p_stats kernel and optimize_p host code are tested against on the GPU.
"""
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


def run_case(name, M, K, D, rng, coupled, unassigned=0, missing=True):
    vm = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar"), os.path.join(REF, "MVTopicModel-1.0-SNAPSHOT.jar")])
    sh = vm.shims
    sh["java/util/ArrayList.size:()I"] = lambda loc, r, a, pc: len(r[1])
    sh["java/util/ArrayList.get:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: r[1][a[0]]
    sh["cc/mallet/types/LabelSequence.getFeatures:()[I"] = lambda loc, r, a, pc: r[1]
    sh["cc/mallet/types/Instance.getData:()Ljava/lang/Object;"] = lambda loc, r, a, pc: ("fs", r[1])
    sh["cc/mallet/types/FeatureSequence.getLength:()I"] = lambda loc, r, a, pc: len(r[1])
    sh["java/lang/Byte.valueOf:(B)Ljava/lang/Byte;"] = lambda loc, r, a, pc: a[0]
    sh["java/lang/Byte.byteValue:()B"] = lambda loc, r, a, pc: r
    for d in ("(Ljava/lang/String;)", "(D)", "(I)", "(Ljava/lang/Object;)"):
        sh["java/lang/StringBuilder.append:" + d + "Ljava/lang/StringBuilder;"] = lambda loc, r, a, pc: r
    sh["java/lang/StringBuilder.toString:()Ljava/lang/String;"] = lambda loc, r, a, pc: ""
    sh["org/apache/log4j/Logger.info:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: None
    sh["java/io/PrintStream.println:(Ljava/lang/String;)V"] = lambda loc, r, a, pc: None
    sh[MC + ".appendMetadata:(Ljava/lang/String;)V"] = lambda loc, r, a, pc: None
    vm.statics[(MC, "logger")] = JObject("logger")
    vm.statics[("java/lang/System", "err")] = ("stderr",)
    # documents with strictly decreasing view lengths; trailing views may be missing
    lens = np.zeros((M, D), dtype=int)
    lens[0] = rng.integers(2 * M + 2, 4 * M + 12, D)
    for m in range(1, M):
        lens[m] = np.maximum(lens[m - 1] - rng.integers(1, 4, D), 1)
        if m == M - 1 and missing:
            lens[m][rng.random(D) < 0.25] = 0
    views, zs, docs = [], [], []
    theta = rng.dirichlet(np.full(K, 0.2), D)
    for m in range(M):
        off = np.concatenate([[0], np.cumsum(lens[m])])
        z = np.concatenate([rng.choice(K, size=lens[m][d], p=theta[d] if coupled else None) for d in range(D)]).astype(int)
        if unassigned:
            z[rng.choice(len(z), unassigned, replace=False)] = -1
        views.append({"off": off.tolist(), "word": [0] * int(off[-1])})
        zs.append(z.tolist())
    for d in range(D):
        ent = JObject("org/madgik/utils/MixTopicModelTopicAssignment")
        asg = []
        for m in range(M):
            b, e = views[m]["off"][d], views[m]["off"][d + 1]
            if e == b:
                asg.append(None); continue
            ta = JObject("cc/mallet/topics/TopicAssignment")
            ta.fields["instance"] = ("instance", [0] * (e - b))
            ta.fields["topicSequence"] = ("labels", zs[m][b:e])
            asg.append(ta)
        ent.fields["Assignments"] = asg
        docs.append(ent)
    docs_per_view = [int(np.sum(lens[m] > 0)) for m in range(M)]
    model = JObject(MC)
    model.fields.update(dict(numModalities=M, numTopics=K, data=("arraylist", docs), totalDocsPerModality=docs_per_view,
                             pMean=[[0.0] * M for _ in range(M)], p_a=[[0.2] * M for _ in range(M)], p_b=[[1.0] * M for _ in range(M)]))
    vm.strict_fields = True
    vm.call(MC, "optimizeP", "(Z)V", [model, 0])
    f = model.fields
    print(f"  {name}: pMean {np.round(np.array(f['pMean']), 4).tolist()} p_a {np.round(np.array(f['p_a']), 3).tolist()} steps {vm.steps}")
    return {"name": name, "M": M, "K": K, "all_views_present": not missing, "views": views, "z": zs, "totalDocsPerModality": docs_per_view,
            "pMean": f["pMean"], "p_a": f["p_a"], "p_b": f["p_b"]}


def main():
    rng = np.random.default_rng(20261018)
    out = {"source": "optimizeP executed from the reference's jar by tools/jvm_mini.py (documents with decreasing view lengths, see the script)",
           "cases": [run_case("two_views_coupled", 2, 8, 60, rng, True),
                     run_case("three_views_independent", 3, 6, 50, rng, False, missing=False),
                     run_case("four_views_unassigned_tokens", 4, 10, 40, rng, True, unassigned=7),
                     run_case("two_views_identical_topics", 2, 1, 20, rng, True, missing=False)]}   # K = 1: pMean = 1 -> a = 5000 -> p_a = 100
    json.dump(out, open(os.path.join(HERE, "reference_optimize_p.json"), "w"))


if __name__ == "__main__":
    main()
