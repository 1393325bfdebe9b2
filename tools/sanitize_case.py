"""Small multi-view + single-view run for compute-sanitizer (memcheck / racecheck): a few sweeps, probe, LL, hist, optimise."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import random_corpus
from mvtopicmodel_b200 import Engine
for K, Vs, means in [(50, [120], [7]), (130, [150, 40, 30], [12, 3, 2]), (600, [200, 60], [20, 4])]:
    views = random_corpus(K, 150, K, Vs, means, oov=True)
    e = Engine(K, Vs, views, seed=3, max_ctas=8, warps_per_cta=4)
    e.init_assignments()
    for it in range(1, 4):
        e.sweep(it)
    e.sweep(4, update_global=False)
    for m in range(len(Vs)):
        e.set_assignments(m, e.get_assignments(m))      # frozen sweep moved z only: rebuild the counts
    off = views[0][0]
    d = int(np.argmax(off[1:] - off[:-1]))
    e.cond_probs(0, d, 0)
    e.loglik(); e.doc_topic_hist(0); e.p_statistics(); e.optimize_hyper(5, 15)
    e.sweep(6)
    assert e.check_invariants() == 0
    # round-2 paths: pipelined host sweep (pageable arrays), async view passes + hand-over, host mirror, held-out scoring
    zs = [e.get_assignments(m).copy() for m in range(len(Vs))]
    e.sweep_host(7, zs)
    import torch
    comm = torch.cuda.Stream()
    e.delta_begin()
    mir = [torch.empty(n, dtype=torch.int32).pin_memory().numpy() for n in e.ntok]
    for m in range(len(Vs)):
        mir[m][:] = e.get_assignments(m); e.set_host_mirror(m, mir[m])
    for m in range(len(Vs)):
        e.sweep_view_async(8, m)
        e.stream_wait_view(m, comm.cuda_stream)
        e.sum_exchange_finish_async(m, 1, comm.cuda_stream, 4)
        e.view_wait_stream(m, comm.cuda_stream)
    e.sweep_finish()
    e.sweep_host(9, mir)                                 # pinned arrays: the kernel writes them
    for m in range(len(Vs)):
        assert np.array_equal(mir[m], e.get_assignments(m))
        e.set_host_mirror(m, None)
    e.heldout_loglik(0, views[0][0], views[0][1])
    assert e.check_invariants() == 0
    e.close()
print("sanitize case done")
