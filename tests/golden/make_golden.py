"""Generates the committed fixtures of tests/golden/ (run in the build container, where /root/reference exists).

  oracle_tiny.json  -- a tiny two-view corpus with the oracle's initial assignments, its assignments after a few
                       reference-faithful sweeps and the resulting log-likelihood (drift guard for the oracle and an
                       integer golden vector for the engine's bit-exact initialisation).
  sms_corpus.npz    -- BASELINE configs[0]: SampleData/SMSSpamCollection2.txt (id \\t label \\t text) through an
                       approximation of the reference's text pipeline (S:1809-1817, S:1843-1845): lower-case, letter
                       tokens, stoplists/en.txt, drop tokens shorter than 3 characters, drop types seen < round(0.001*D)
                       times.  MALLET's SimpleTokenizer is binary-only, so the vocabulary is approximate.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_oracle_tiny():
    from helpers import random_corpus
    from oracle import oracle as O
    K, Vs = 12, [30, 10]
    views = random_corpus(101, 40, K, Vs, [6, 2])
    o = O.Oracle(K, Vs, views, seed=424242)
    o.init_assignments()
    z_init = [o.get_assignments(m).tolist() for m in range(2)]
    flags, sweeps = O.F_STALE_TREES, 5
    for it in range(1, sweeps + 1):
        o.sweep(it, flags)
    g = {"K": K, "V": Vs, "seed": 424242, "flags": flags, "sweeps": sweeps,
         "views": [{"off": v[0].tolist(), "word": v[1].tolist()} for v in views],
         "z_init": z_init, "z_final": [o.get_assignments(m).tolist() for m in range(2)], "loglik": o.loglik().tolist()}
    json.dump(g, open(os.path.join(HERE, "oracle_tiny.json"), "w"))
    print("oracle_tiny.json: tokens", [len(z) for z in z_init], "loglik", g["loglik"])


def make_sms():
    ref = "/root/reference"
    path = os.path.join(ref, "SampleData", "SMSSpamCollection2.txt")
    stop = set(w.strip().lower() for w in open(os.path.join(ref, "stoplists", "en.txt"), encoding="utf-8", errors="ignore") if w.strip())
    docs = []
    for line in open(path, encoding="utf-8", errors="ignore"):
        parts = line.rstrip("\n").split("\t", 2)
        if len(parts) < 3:
            continue
        toks = [t for t in re.findall(r"[^\W\d_]+", parts[2].lower()) if len(t) >= 3 and t not in stop]
        docs.append(toks)
    D = len(docs)
    from collections import Counter
    cnt = Counter(t for d in docs for t in d)
    prune = int(round(0.001 * D))
    vocab = sorted(t for t, c in cnt.items() if c >= prune)
    idx = {t: i for i, t in enumerate(vocab)}
    off = np.zeros(D + 1, dtype=np.int64)
    words = []
    for d, toks in enumerate(docs):
        ids = [idx[t] for t in toks if t in idx]
        words.extend(ids)
        off[d + 1] = len(words)
    np.savez_compressed(os.path.join(HERE, "sms_corpus.npz"), doc_off=off, word_id=np.array(words, dtype=np.int32), V=np.int32(len(vocab)))
    print("sms_corpus.npz: docs", D, "types", len(vocab), "tokens", len(words), "max len", int((off[1:] - off[:-1]).max()))


if __name__ == "__main__":
    make_oracle_tiny()
    make_sms()
