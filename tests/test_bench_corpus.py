"""bench.py's corpus: every --gpus N samples the SAME corpus (strong scaling) -- the union of 8 blocks; rank r of N holds blocks
r, r+N, ...  CPU test of the host logic (no device)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _docs(views):
    """documents as tuples of per-view word tuples, in order"""
    D = len(views[0][0]) - 1
    return [tuple(tuple(w[off[d]:off[d + 1]].tolist()) for off, w in views) for d in range(D)]


def test_ranks_partition_one_corpus():
    import bench
    total = 8 * 40
    K1, V1, whole = bench.build_corpus("tiny_2v", total, 0, 1)
    docs1 = _docs(whole)
    assert len(docs1) == total and K1 == 20 and list(V1) == [200, 80]
    per = total // bench.N_BLOCKS
    blocks1 = [docs1[b * per:(b + 1) * per] for b in range(bench.N_BLOCKS)]
    for world in (2, 4, 8):
        seen = 0
        for rank in range(world):
            K, V, views = bench.build_corpus("tiny_2v", total, rank, world)
            docs = _docs(views)
            mine = list(range(rank, bench.N_BLOCKS, world))
            assert len(docs) == per * len(mine)
            for j, b in enumerate(mine):                      # block b of the one-GPU corpus, document for document
                assert docs[j * per:(j + 1) * per] == blocks1[b], (world, rank, b)
            seen += len(docs)
        assert seen == total
    # deterministic: a second build gives the same arrays
    K2, V2, again = bench.build_corpus("tiny_2v", total, 0, 1)
    for (o1, w1), (o2, w2) in zip(whole, again):
        assert np.array_equal(o1, o2) and np.array_equal(w1, w2)


def test_roofline_bytes_per_token_formula():
    import bench
    # SURVEY 8(d) / BASELINE.md section 3: 4K + 28 + 8K/N (+ 4 * other views' lengths / N for multi-view)
    assert bench.b_tok_view(500, 0, [200.0]) == 4 * 500 + 28 + 8 * 500 / 200
    assert abs(bench.b_tok_view(1000, 0, [120.0, 6.0]) - (4000 + 28 + 8000 / 120 + 4 * 6 / 120)) < 1e-9
    assert abs(bench.b_tok(1000, [120.0, 6.0], [120, 6]) - (120 * bench.b_tok_view(1000, 0, [120.0, 6.0]) + 6 * bench.b_tok_view(1000, 1, [120.0, 6.0])) / 126) < 1e-9
