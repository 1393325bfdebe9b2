"""getSortedWords(modality) (M:1792-1809) EXECUTED from the reference's jar, with java.util.TreeSet replaced by a sorted list whose
ordering is decided by MALLET's own IDSorter.compareTo bytecode (output/lib/mallet-2.0.8.jar) -- so the order of equal counts
(larger type id first) comes from the reference's binaries, not from a reading of them.

Output: tests/golden/reference_sorted_words.json; tests/test_ingest_state.py compares state_io.sorted_words / top_words with it.
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
MC, IDS = "org/madgik/MVTopicModel/FastQMVWVParallelTopicModel", "cc/mallet/types/IDSorter"


def main():
    vm = jvm_mini.MiniJVM([os.path.join(REF, "lib", "mallet-2.0.8.jar"), os.path.join(REF, "MVTopicModel-1.0-SNAPSHOT.jar")])
    sh = vm.shims

    def list_init(loc, r, a, pc):
        r.fields["items"] = []

    def list_add(loc, r, a, pc):
        r.fields["items"].append(a[0]); return 1

    def tree_add(loc, r, a, pc):                         # TreeSet.add: binary search by compareTo; an equal element is not added
        items, x = r.fields["items"], a[0]
        lo, hi = 0, len(items)
        while lo < hi:
            mid = (lo + hi) // 2
            cmp = vm.call(IDS, "compareTo", "(Lcc/mallet/types/IDSorter;)I", [x, items[mid]])
            if cmp == 0:
                return 0
            if cmp < 0:
                hi = mid
            else:
                lo = mid + 1
        items.insert(lo, x)
        return 1
    sh["java/util/ArrayList.<init>:(I)V"] = list_init
    sh["java/util/ArrayList.add:(Ljava/lang/Object;)Z"] = list_add
    sh["java/util/TreeSet.<init>:()V"] = list_init
    sh["java/util/TreeSet.add:(Ljava/lang/Object;)Z"] = tree_add
    rng = np.random.default_rng(20261018)
    out = {"source": "getSortedWords executed from the reference's jar; ordering by MALLET's IDSorter.compareTo bytecode", "cases": []}
    for V, K, hi in [(40, 5, 4), (200, 7, 3), (12, 3, 50)]:          # small count ranges: many ties
        nwk = (rng.integers(0, hi, size=(V, K)) * (rng.random((V, K)) < 0.6)).astype(int)
        nwk[:, K - 1] = 0 if V == 12 else nwk[:, K - 1]              # a topic without words
        model = JObject(MC)
        model.fields.update(dict(numTopics=K, numTypes=[7, V], typeTopicCounts=[None, nwk.tolist()]))
        res = vm.call(MC, "getSortedWords", "(I)Ljava/util/ArrayList;", [model, 1])
        srt = [[[o.fields["id"], int(o.fields["p"])] for o in ts.fields["items"]] for ts in res.fields["items"]]
        out["cases"].append({"V": V, "K": K, "typeTopicCounts": nwk.tolist(), "sorted": srt})
    # displayTopWords(numWords, numLabels, usingNewLines) (M:1851-1888) from the same jar: real string building, numbers through the
    # mirror's NumberFormat restatement, words as their type ids.  The jar predates the source in ONE character: its one-line mode
    # separates words by " " where the source (and the mirror) writes "; " -- recorded as is.
    sys.path.insert(0, ROOT)
    from mvtopicmodel_b200 import state_io

    def sb_init(loc, r, a, pc):
        r.fields["s"] = ""

    def sb_append(loc, r, a, pc):
        v = a[0]
        r.fields["s"] += v if isinstance(v, str) else str(v)
        return r
    sh["java/lang/StringBuilder.<init>:()V"] = sb_init
    for d in ("(I)", "(Ljava/lang/String;)", "(Ljava/lang/Object;)"):
        sh["java/lang/StringBuilder.append:" + d + "Ljava/lang/StringBuilder;"] = sb_append
    sh["java/lang/StringBuilder.toString:()Ljava/lang/String;"] = lambda loc, r, a, pc: r.fields["s"]
    sh["java/text/NumberFormat.format:(D)Ljava/lang/String;"] = lambda loc, r, a, pc: state_io.java_number_format(a[0])
    sh["cc/mallet/types/Alphabet.lookupObject:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: str(a[0])
    sh["java/util/ArrayList.get:(I)Ljava/lang/Object;"] = lambda loc, r, a, pc: r.fields["items"][a[0]]
    sh["java/lang/Byte.valueOf:(B)Ljava/lang/Byte;"] = lambda loc, r, a, pc: a[0]
    sh["java/lang/Byte.byteValue:()B"] = lambda loc, r, a, pc: r

    class _It:
        def __init__(self, lst):
            self.l, self.i = lst, 0
    sh["java/util/TreeSet.iterator:()Ljava/util/Iterator;"] = lambda loc, r, a, pc: _It(r.fields["items"])
    sh["java/util/Iterator.hasNext:()Z"] = lambda loc, r, a, pc: int(r.i < len(r.l))

    def _next(loc, r, a, pc):
        v = r.l[r.i]; r.i += 1
        return v
    sh["java/util/Iterator.next:()Ljava/lang/Object;"] = _next
    out["displayTopWords"] = []
    V, K = 30, 4
    nwk = [(rng.integers(0, 6, size=(V, K)) * (rng.random((V, K)) < 0.5)).astype(int) for _ in range(2)]
    alpha = rng.dirichlet(np.full(K + 1, 0.8), size=2)
    model = JObject(MC)
    model.fields.update(dict(numTopics=K, numModalities=2, numTypes=[V, V], typeTopicCounts=[t.tolist() for t in nwk],
                             alpha=alpha.tolist(), formatter=("nf",), alphabet=[("alphabet", 0), ("alphabet", 1)]))
    for num_words, new_lines in [(1, 0), (4, 0), (4, 1), (100, 1)]:
        text = vm.call(MC, "displayTopWords", "(IIZ)Ljava/lang/String;", [model, num_words, 0, new_lines])
        out["displayTopWords"].append({"numWords": num_words, "usingNewLines": bool(new_lines), "text": text})
    out["displayTopWords_state"] = {"typeTopicCounts": [t.tolist() for t in nwk], "alpha": alpha.tolist()}
    json.dump(out, open(os.path.join(HERE, "reference_sorted_words.json"), "w"))
    print("reference_sorted_words.json:", [(c["V"], c["K"], [len(s) for s in c["sorted"]]) for c in out["cases"]], "bytecode steps", vm.steps)


if __name__ == "__main__":
    main()
