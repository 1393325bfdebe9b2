"""cc.mallet.pipe.SimpleTokenizer.pipe EXECUTED from output/lib/mallet-2.0.8.jar (tools/jvm_mini.py) over the lines of
SampleData/SMSSpamCollection2.txt (lower-cased, as CharSequenceLowercase does before it in the reference's pipe list S:1809-1817)
with stoplists/en.txt as its stop list.  Shimmed: java.lang.Character (code points from the Python string; getType through the
Unicode general category, the same table Java uses), String(int[], int, int), the HashSet / ArrayList / Instance containers.
Output: tests/golden/reference_tokenizer_vectors.json -- a checksum of the token stream of every line plus the full token lists
of the first 300 lines and of a few hand-made edge cases; tests/test_ingest_state.py holds mvtopicmodel_b200.ingest.simple_tokenize
to them.  Needs /root/reference."""
import hashlib
import json
import os
import sys
import unicodedata

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import jvm_mini  # noqa: E402
from mvtopicmodel_b200 import ingest  # noqa: E402

REF = "/root/reference"
ST = "cc/mallet/pipe/SimpleTokenizer"
JAVA_TYPE = ingest._JAVA_TYPE

EDGE = ["abc123def", "abc\tdef\nghi", "don't stop", "snake_case-word", "a+b=c $5 ok", "café não", "“quoted” (paren) [x]", "",
        "123 456", "x" * 2500, "éclair with combining mark", "MiXeD Case STAYS as given", "tab\tseparated\tvalues",
        "semi;colon:colon,comma.period!bang?", "under__score __lead trail__", "中文 mixed with latin", "emoji \U0001F600 face"]


def tokenizer_vm(stoplist):
    vm = jvm_mini.MiniJVM([os.path.join(REF, "output", "lib", "mallet-2.0.8.jar")])
    sh = vm.shims
    sh["java/lang/Character.codePointAt:(Ljava/lang/CharSequence;I)I"] = lambda loc, r, a, pc: ord(a[0][a[1]])
    sh["java/lang/Character.codePointCount:(Ljava/lang/CharSequence;II)I"] = lambda loc, r, a, pc: a[2] - a[1]
    sh["java/lang/Character.getType:(I)I"] = lambda loc, r, a, pc: JAVA_TYPE.get(unicodedata.category(chr(a[0])), 0)
    sh["java/lang/CharSequence.length:()I"] = lambda loc, r, a, pc: len(r)

    def string_init(loc, r, a, pc):                      # new String(int[] codePoints, int offset, int count)
        r.fields["value"] = "".join(chr(c) for c in a[0][a[1]:a[1] + a[2]])
    sh["java/lang/String.<init>:([III)V"] = string_init
    sh["java/util/HashSet.contains:(Ljava/lang/Object;)Z"] = lambda loc, r, a, pc: int(a[0].fields["value"] in stoplist)

    def list_add(loc, r, a, pc):
        r.fields.setdefault("items", []).append(a[0].fields["value"]); return 1
    sh["java/util/ArrayList.add:(Ljava/lang/Object;)Z"] = list_add
    sh["cc/mallet/types/Instance.getData:()Ljava/lang/Object;"] = lambda loc, r, a, pc: r["data"]

    def set_data(loc, r, a, pc):
        r["data"] = a[0]
    sh["cc/mallet/types/Instance.setData:(Ljava/lang/Object;)V"] = set_data
    tok = jvm_mini.JObject(ST)
    tok.fields["stoplist"] = ("stoplist",)
    return vm, tok


def run(vm, tok, text):
    # Java strings are UTF-16: the bytecode indexes chars; keep to the BMP or pre-split astral characters as Java would see them
    inst = {"data": text}
    vm.call(ST, "pipe", "(Lcc/mallet/types/Instance;)Lcc/mallet/types/Instance;", [tok, inst])
    return list(inst["data"].fields.get("items", []))


def main():
    stop = ingest.load_stoplist(os.path.join(REF, "stoplists", "en.txt"))
    vm, tok = tokenizer_vm(stop)
    docs = ingest.read_sms_collection(os.path.join(REF, "SampleData", "SMSSpamCollection2.txt"))
    h = hashlib.sha256()
    first, n_tok = [], 0
    for i, (_, text) in enumerate(docs):
        low = text.lower()
        if any(ord(c) > 0xFFFF for c in low):
            low = "".join(c for c in low if ord(c) <= 0xFFFF)           # (none in this file; see `run`)
        toks = run(vm, tok, low)
        n_tok += len(toks)
        h.update(("\x1f".join(toks) + "\x1e").encode("utf-8"))
        if i < 300:
            first.append(toks)
    vm2, tok2 = tokenizer_vm(set())                      # the edge cases run with an empty stop list (no reference data needed to check them)
    edge = [[t, run(vm2, tok2, t)] for t in EDGE if all(ord(c) <= 0xFFFF for c in t)]
    out = {"source": "cc.mallet.pipe.SimpleTokenizer.pipe executed from mallet-2.0.8.jar by tools/jvm_mini.py",
           "lines": len(docs), "tokens": n_tok, "sha256_of_token_stream": h.hexdigest(), "first_lines": first, "edge_cases": edge}
    json.dump(out, open(os.path.join(HERE, "reference_tokenizer_vectors.json"), "w"))
    print("lines", len(docs), "tokens", n_tok, "sha256", h.hexdigest()[:16], "bytecode steps", vm.steps)


if __name__ == "__main__":
    main()
