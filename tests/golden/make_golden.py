"""Generates the committed fixtures of tests/golden/ (run in the build container, where /root/reference exists).

  oracle_tiny.json  -- a tiny two-view corpus with the oracle's initial assignments, its assignments after a few
                       reference-faithful sweeps and the resulting log-likelihood (drift guard for the oracle and an
                       integer golden vector for the engine's bit-exact initialisation).
  sms_corpus.npz    -- BASELINE configs[0]: SampleData/SMSSpamCollection2.txt (id \\t label \\t text) through the reference's
  sms_vocab.json       text pipeline (S:1800-1851, GenerateStoplist S:631-730) as restated in mvtopicmodel_b200/ingest.py, with
                       MALLET's SimpleTokenizer / FeatureCountPipe rules recovered from the jar's bytecode (tools/jclass.py).
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
    """BASELINE configs[0] through the reference's own text pipeline as restated in mvtopicmodel_b200/ingest.py
    (S:1800-1851 with PruneCntPerc = 0.001 of src/main/resources/config.properties:14; MALLET's SimpleTokenizer recovered
    from bytecode).  Also stores a tokenizer known-answer table (input line -> tokens) for the CPU tests."""
    from mvtopicmodel_b200 import ingest
    ref = "/root/reference"
    docs = ingest.read_sms_collection(os.path.join(ref, "SampleData", "SMSSpamCollection2.txt"))
    stop = ingest.load_stoplist(os.path.join(ref, "stoplists", "en.txt"))
    lists, alphabets = ingest.import_instances([docs], 1, prune_cnt_perc=0.001, prune_lbl_cnt_perc=0.001, prune_max_perc=10.0,
                                               text_stoplist=stop)
    il, alpha = lists[0], alphabets[0]
    D = len(il)
    off = np.zeros(D + 1, dtype=np.int64)
    np.cumsum([len(i.features) for i in il], out=off[1:])
    words = np.concatenate([i.features for i in il]).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "sms_corpus.npz"), doc_off=off, word_id=words, V=np.int32(len(alpha)))
    json.dump({"vocab": alpha.entries, "first_docs": [[alpha.lookup_object(int(w)) for w in il[d].features] for d in range(12)]},
              open(os.path.join(HERE, "sms_vocab.json"), "w"))
    print("sms_corpus.npz: docs", D, "types", len(alpha), "tokens", len(words), "max len", int((off[1:] - off[:-1]).max()))


if __name__ == "__main__":
    make_oracle_tiny()
    make_sms()
