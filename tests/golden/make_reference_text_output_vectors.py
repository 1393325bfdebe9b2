"""printState(PrintStream) (M:3276-3320), printTypeTopicCounts(File) (M:2076-2102) and printTopicWordWeights(PrintWriter)
(M:2113-2129) EXECUTED from the reference's jar by tools/jvm_mini.py over a small two-view state.  The JDK classes the methods
write through are replaced by string collectors (PrintStream / PrintWriter / StringBuilder / Formatter("%d %s %d %d %s %d\\n"));
Double.toString is the mirror's restatement (state_io.java_double_to_string, tested on its own against known JDK outputs), so what
these vectors pin is everything else: line order, separators, which views and tokens are written, zero counts included or not.

Output: tests/golden/reference_text_outputs.json; tests/test_ingest_state.py compares state_io.write_state /
write_type_topic_counts / write_topic_word_weights with it.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
import jvm_mini  # noqa: E402
from jvm_mini import JObject  # noqa: E402
from mvtopicmodel_b200 import state_io  # noqa: E402

REF = "/root/reference/output"
MC = "org/madgik/MVTopicModel/FastQMVWVParallelTopicModel"


def main():
    vm = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar"), os.path.join(REF, "MVTopicModel-1.0-SNAPSHOT.jar")])
    sh = vm.shims
    sink = []                                            # everything the method prints, in order

    def jstr(v):
        if isinstance(v, JObject) and "s" in v.fields:
            return v.fields["s"]
        if isinstance(v, float):
            return state_io.java_double_to_string(v)
        return v if isinstance(v, str) else str(v)

    def sb_init(loc, r, a, pc):
        r.fields["s"] = ""

    def sb_append(loc, r, a, pc):
        r.fields["s"] += jstr(a[0]); return r
    sh["java/lang/StringBuilder.<init>:()V"] = sb_init
    for d in ("(I)", "(D)", "(Ljava/lang/String;)", "(Ljava/lang/Object;)"):
        sh["java/lang/StringBuilder.append:" + d + "Ljava/lang/StringBuilder;"] = sb_append
    sh["java/lang/StringBuilder.toString:()Ljava/lang/String;"] = lambda loc, r, a, pc: r.fields["s"]
    for cls in ("java/io/PrintStream", "java/io/PrintWriter"):
        sh[cls + ".println:(Ljava/lang/String;)V"] = lambda loc, r, a, pc: sink.append(jstr(a[0]) + "\n")
        sh[cls + ".println:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: sink.append(jstr(a[0]) + "\n")
        sh[cls + ".println:()V"] = lambda loc, r, a, pc: sink.append("\n")
        sh[cls + ".print:(Ljava/lang/String;)V"] = lambda loc, r, a, pc: sink.append(jstr(a[0]))
        sh[cls + ".print:(Ljava/lang/Object;)V"] = lambda loc, r, a, pc: sink.append(jstr(a[0]))
        sh[cls + ".close:()V"] = lambda loc, r, a, pc: None
    sh["java/io/FileWriter.<init>:(Ljava/io/File;)V"] = lambda loc, r, a, pc: None
    sh["java/io/PrintWriter.<init>:(Ljava/io/Writer;)V"] = lambda loc, r, a, pc: None

    def fmt_init(loc, r, a, pc):
        r.fields["s"] = ""

    def fmt_format(loc, r, a, pc):
        assert a[0] == "%d %s %d %d %s %d\n", a[0]
        r.fields["s"] += a[0] % tuple(jstr(x) if not isinstance(x, int) else x for x in a[1]); return r
    sh["java/util/Formatter.<init>:(Ljava/lang/Appendable;Ljava/util/Locale;)V"] = fmt_init
    sh["java/util/Formatter.format:(Ljava/lang/String;[Ljava/lang/Object;)Ljava/util/Formatter;"] = fmt_format
    vm.statics[("java/util/Locale", "US")] = ("locale",)
    sh["java/lang/Byte.valueOf:(B)Ljava/lang/Byte;"] = lambda loc, r, a, pc: a[0]
    sh["java/lang/Byte.byteValue:()B"] = lambda loc, r, a, pc: r
    sh["java/lang/Integer.valueOf:(I)Ljava/lang/Integer;"] = lambda loc, r, a, pc: a[0]
    sh["java/util/ArrayList.size:()I"] = lambda loc, r, a, pc: len(r[1])
    sh["java/util/ArrayList.get:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: r[1][a[0]]
    sh["cc/mallet/types/Instance.getData:()Ljava/lang/Object;"] = lambda loc, r, a, pc: ("fs", r[1])
    sh["cc/mallet/types/Instance.getSource:()Ljava/lang/Object;"] = lambda loc, r, a, pc: r[2]
    sh["java/lang/Object.toString:()Ljava/lang/String;"] = lambda loc, r, a, pc: jstr(r)
    sh["cc/mallet/types/LabelSequence.getLength:()I"] = lambda loc, r, a, pc: len(r[1])
    sh["cc/mallet/types/LabelSequence.getIndexAtPosition:(I)I"] = lambda loc, r, a, pc: r[1][a[0]]
    sh["cc/mallet/types/FeatureSequence.getIndexAtPosition:(I)I"] = lambda loc, r, a, pc: r[1][a[0]]
    sh["cc/mallet/types/Alphabet.lookupObject:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: r[1][a[0]]

    rng = np.random.default_rng(20261018)
    K, Vs, D = 4, [9, 5], 6
    vocab = [[f"w{m}_{i}" for i in range(V)] for m, V in enumerate(Vs)]
    vocab[0][3] = "Deep Learning"                        # a type with a blank (keyphrase views have them)
    lens = [rng.integers(1, 6, D), rng.integers(1, 4, D)]
    views, zs, docs = [], [], []
    for m in range(2):
        off = np.concatenate([[0], np.cumsum(lens[m])])
        views.append((off.tolist(), rng.integers(0, Vs[m], int(off[-1])).tolist()))
        zs.append(rng.integers(0, K, int(off[-1])).tolist())
    sources = [None, "src1", "a/b.txt", None, "id:77", "x"]
    for d in range(D):
        ent = JObject("org/madgik/utils/MixTopicModelTopicAssignment")
        asg = []
        for m in range(2):
            b, e = views[m][0][d], views[m][0][d + 1]
            ta = JObject("cc/mallet/topics/TopicAssignment")
            ta.fields["instance"] = ("instance", views[m][1][b:e], sources[d])
            ta.fields["topicSequence"] = ("labels", zs[m][b:e])
            asg.append(ta)
        ent.fields["Assignments"] = asg
        docs.append(ent)
    nwk = [np.zeros((V, K), dtype=int) for V in Vs]
    for m in range(2):
        for w, t in zip(views[m][1], zs[m]):
            nwk[m][w, t] += 1
    alpha = rng.dirichlet(np.full(K + 1, 1.0), size=2)
    gamma, beta = [1.25, 0.5], [0.01, 0.123456789]
    model = JObject(MC)
    model.fields.update(dict(numTopics=K, numModalities=2, numTypes=list(Vs), typeTopicCounts=[t.tolist() for t in nwk],
                             alpha=alpha.tolist(), gamma=gamma, beta=beta, data=("arraylist", docs),
                             alphabet=[("alphabet", vocab[0]), ("alphabet", vocab[1])]))
    out = {"source": "printState / printTypeTopicCounts / printTopicWordWeights executed from the reference's jar by tools/jvm_mini.py",
           "K": K, "V": Vs, "vocab": vocab, "views": [{"off": v[0], "word": v[1]} for v in views], "z": zs, "sources": sources,
           "alpha": alpha.tolist(), "gamma": gamma, "beta": beta, "typeTopicCounts": [t.tolist() for t in nwk]}
    vm.call(MC, "printState", "(Ljava/io/PrintStream;)V", [model, ("printstream",)])
    out["printState"] = "".join(sink); sink.clear()
    vm.call(MC, "printTypeTopicCounts", "(Ljava/io/File;)V", [model, ("file",)])
    out["printTypeTopicCounts"] = "".join(sink); sink.clear()
    vm.call(MC, "printTopicWordWeights", "(Ljava/io/PrintWriter;)V", [model, ("printwriter",)])
    out["printTopicWordWeights"] = "".join(sink); sink.clear()
    json.dump(out, open(os.path.join(HERE, "reference_text_outputs.json"), "w"))
    print("reference_text_outputs.json:", {k: len(out[k]) for k in ("printState", "printTypeTopicCounts", "printTopicWordWeights")}, "chars; bytecode steps", vm.steps)
    print(out["printState"][:400])


if __name__ == "__main__":
    main()
