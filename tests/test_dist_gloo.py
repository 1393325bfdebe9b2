"""CPU tests of the multi-rank host logic (world_size 2, gloo): document sharding with global ids, and the integer
delta all-reduce protocol of mvtopicmodel_b200/dist.py.  The per-rank sampler here is the CPU oracle (a test
stand-in for the CUDA engine, which needs a GPU); the protocol code under test is the product's CountExchange."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleAdapter:
    """delta_begin / delta_reset / delta_export / delta_import over an Oracle, mirroring EngineAdapter on CPU tensors."""

    def __init__(self, o):
        self.o, self.M = o, o.M
        self.snap = [None] * o.M
        self.buf = [None] * o.M

    def _cur(self, m):
        nwk, nk = self.o.get_counts(m)
        return torch.from_numpy(nwk.reshape(-1).copy()), torch.from_numpy(nk.copy())

    def delta_begin(self):
        self.snap = [self._cur(m) for m in range(self.M)]

    def delta_reset(self):
        self.snap = [tuple(torch.zeros_like(t) for t in self._cur(m)) for m in range(self.M)]

    def delta_export(self, m):
        a, b = self._cur(m)
        self.buf[m] = (a - self.snap[m][0], b - self.snap[m][1])
        return self.buf[m]

    def sum_buffers(self, m):
        self.buf[m] = self._cur(m)
        return self.buf[m]

    def sum_finish(self, m, world):
        a = self.buf[m][0] - (world - 1) * self.snap[m][0]
        b = self.buf[m][1] - (world - 1) * self.snap[m][1]
        self.snap[m] = (a, b)
        self.o.set_counts(m, a.numpy().reshape(int(self.o.V[m]), self.o.K), b.numpy())

    def delta_import(self, m):
        a = self.snap[m][0] + self.buf[m][0]
        b = self.snap[m][1] + self.buf[m][1]
        self.snap[m] = (a, b)
        self.o.set_counts(m, a.numpy().reshape(int(self.o.V[m]), self.o.K), b.numpy())


class OracleOverlapAdapter:
    """The adapter surface OverlappedSweep drives (OverlapAdapter on a GPU), over the CPU oracle: the oracle sweeps all views
    of a document together, so the whole sweep runs when view 0 is queued; streams collapse to program order."""

    def __init__(self, base):
        self.b, self.M, self.log = base, base.M, []
        self.whole = [None] * base.M

    def sweep_view_async(self, it, m):
        self.log.append(("pass", m))
        if m == 0:
            from oracle import oracle as O
            self.b.o.sweep(it, O.F_ENGINE_MIRROR)

    def comm_wait_view(self, m):
        self.log.append(("comm_waits_pass", m))

    def whole_buffer(self, m):
        a, b = self.b._cur(m)
        self.whole[m] = torch.cat([a, b])
        return self.whole[m]

    def comm_context(self):
        import contextlib
        return contextlib.nullcontext()

    def finish_async(self, m, world):
        n = self.whole[m].numel() - self.b.o.K
        self.b.buf[m] = (self.whole[m][:n], self.whole[m][n:])
        self.b.sum_finish(m, world)
        self.log.append(("finish", m))

    def view_wait_comm(self, m):
        self.log.append(("pass_waits_comm", m))

    def sweep_finish(self):
        self.log.append(("barrier", -1))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import random_corpus, recount
        from mvtopicmodel_b200 import corpus
        from mvtopicmodel_b200.dist import CountExchange, shard_doc_ids
        from oracle import oracle as O
        K, Vs = 24, [80, 30]
        full = random_corpus(17, 301, K, Vs, [9, 3])
        D = len(full[0][0]) - 1
        base, stride, n_local = shard_doc_ids(D, rank, world)
        views = corpus.shard_views(full, rank, world)
        assert len(views[0][0]) - 1 == n_local
        o = O.Oracle(K, Vs, views, seed=5)
        o.set_doc_ids(base, stride)
        x = CountExchange(OracleAdapter(o))
        x.reset()
        o.init_assignments()
        # (1) a sharded run draws what the unsharded one does: compare with the unsharded oracle's init
        ref = O.Oracle(K, Vs, full, seed=5)
        ref.init_assignments()
        for m in range(2):
            zr, zl = ref.get_assignments(m), o.get_assignments(m)
            off = full[m][0]
            mine = np.concatenate([zr[off[d]:off[d + 1]] for d in range(rank, D, world)]) if len(zl) else zl
            assert np.array_equal(mine, zl)
        x.exchange()
        for m in range(2):      # (2) after the exchange every rank holds the global counts of the unsharded init
            a, b = o.get_counts(m); ra, rb = ref.get_counts(m)
            assert np.array_equal(a, ra) and np.array_equal(b, rb)
        # (3) sweeps with the exchange keep the global invariants bit-exact
        for it in range(1, 5):
            o.sweep(it, O.F_ENGINE_MIRROR)
            nbytes = x.exchange() if it % 2 else x.exchange_sum()      # both forms of the protocol
            assert nbytes == sum((Vs[m] * K + K) * 4 for m in range(2))
            zs_all = [None] * world
            dist.all_gather_object(zs_all, [o.get_assignments(m) for m in range(2)])
            for m in range(2):
                nwk = np.zeros((Vs[m], K), dtype=np.int64); nk = np.zeros(K, dtype=np.int64)
                for r in range(world):
                    vr = corpus.shard_views(full, r, world)
                    (a, b), = recount([vr[m]], [zs_all[r][m]], K, [Vs[m]])
                    nwk += a; nk += b
                a, b = o.get_counts(m)
                assert np.array_equal(a, nwk) and np.array_equal(b, nk)
        # (4) the overlapped form (one all-reduce per view over table + totals, hand-over calls in protocol order)
        from mvtopicmodel_b200.dist import OverlappedSweep
        oa = OracleOverlapAdapter(x.a)
        ovl = OverlappedSweep(oa)
        for it in range(5, 7):
            assert ovl.step(it) == sum((Vs[m] * K + K) * 4 for m in range(2))
            zs_all = [None] * world
            dist.all_gather_object(zs_all, [o.get_assignments(m) for m in range(2)])
            for m in range(2):
                nwk = np.zeros((Vs[m], K), dtype=np.int64); nk = np.zeros(K, dtype=np.int64)
                for r in range(world):
                    vr = corpus.shard_views(full, r, world)
                    (a, b), = recount([vr[m]], [zs_all[r][m]], K, [Vs[m]])
                    nwk += a; nk += b
                a, b = o.get_counts(m)
                assert np.array_equal(a, nwk) and np.array_equal(b, nk)
        per_view = [("pass", 0), ("comm_waits_pass", 0), ("finish", 0), ("pass_waits_comm", 0)]
        assert oa.log[:4] == per_view and oa.log[8] == ("barrier", -1) and len(oa.log) == 18
        q.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_delta_exchange_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_views_partition():
    from helpers import random_corpus
    from mvtopicmodel_b200 import corpus
    full = random_corpus(3, 50, 8, [20, 10], [5, 2])
    for world in (1, 2, 4, 8):
        tot = [0, 0]
        for r in range(world):
            sv = corpus.shard_views(full, r, world)
            for m in range(2):
                tot[m] += len(sv[m][1])
                assert sv[m][0][-1] == len(sv[m][1])
        assert tot == [len(full[0][1]), len(full[1][1])]


def _hyper_worker(rank, world, port, q):
    """Multi-rank hyper-parameter step on the CPU: each rank holds the statistics of ITS documents (topicDocCounts with its own
    stride, docLengthCounts with its own length); the product's reducer (dist.make_stat_reducer, what mvtm_set_stat_reducer
    installs) makes them global -- max for the strides, sum for the bins -- and the product's hyper core
    (mvtm_test_hyper_core) then computes, on every rank, exactly what the reference's bytecode computed on the whole corpus."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        import json
        from mvtopicmodel_b200 import _lib
        from mvtopicmodel_b200.dist import make_stat_reducer
        lib = _lib.lib()
        reduce_ = make_stat_reducer()
        case = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_hyper_step_vectors.json")))["cases"][1]
        M, K = case["M"], case["K"]
        rng = np.random.default_rng(99)                      # same stream on both ranks: a consistent split of every bin
        hist, lencnt = [], []
        for m in range(M):
            h = np.array(case["topicDocCounts"][m], dtype=np.int64)
            part = rng.binomial(h, 0.5)
            mine = part if rank == 0 else h - part
            # local stride = up to the last non-empty bin of THIS rank, as doc_topic_hist_host sees it before the reduction
            nz = np.nonzero(mine.sum(axis=0))[0]
            ls = np.array([int(nz[-1]) + 1 if len(nz) else 1], dtype=np.int64)
            gs = ls.copy(); reduce_(1, gs, None)               # global max
            loc = np.zeros((K, int(gs[0])), dtype=np.int64); loc[:, :int(ls[0])] = mine[:, :int(ls[0])]
            flat = loc.reshape(-1); reduce_(0, flat, None)      # global sum
            g = flat.reshape(K, int(gs[0]))
            assert np.array_equal(g, h[:, :int(gs[0])]) and not h[:, int(gs[0]):].any()
            hist.append(np.ascontiguousarray(g))
            l = np.array(case["docLengthCounts"][m], dtype=np.int64)
            lp = rng.binomial(l, 0.5)
            lm = lp if rank == 0 else l - lp
            reduce_(0, lm, None)
            assert np.array_equal(lm, l)
            lencnt.append(lm)
        stride = np.array([h.shape[1] for h in hist], dtype=np.int32)
        n_len = np.array([len(l) for l in lencnt], dtype=np.int32)
        hp = (C.c_void_p * M)(*[h.ctypes.data for h in hist]); lp_ = (C.c_void_p * M)(*[l.ctypes.data for l in lencnt])
        alpha = np.array(case["in"]["alpha"]); asum = alpha.sum(axis=1).copy(); gamma = np.array(case["in"]["gamma"])
        gview = np.array(case["in"]["gammaView"]); tables = np.zeros(M); scal = np.array([case["in"]["gammaRoot"], 0.0])
        vals = np.ascontiguousarray([s[3] for s in case["script"]], dtype=np.float64)
        inact, n_inact, used = np.zeros(K, dtype=np.int32), C.c_int32(-1), C.c_int64(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib.mvtm_test_hyper_core(M, K, _lib.OPT_DP | _lib.OPT_GAMMA, hp, p(stride), lp_, p(n_len), p(alpha), p(asum), p(gamma), p(gview),
                                      p(tables), p(scal), p(inact), C.byref(n_inact), p(vals), len(vals), None, C.byref(used))
        assert rc == 0 and used.value == len(vals)
        assert np.allclose(alpha, np.array(case["after_optimizeDP"]["alpha"]), rtol=1e-12, atol=1e-300)
        assert np.allclose(gamma, case["after_optimizeGamma"]["gamma"], rtol=1e-12)
        assert scal[0] == pytest.approx(case["after_optimizeGamma"]["gammaRoot"], rel=1e-12)
        assert inact[:n_inact.value].tolist() == case["after_optimizeDP"]["inactive"]
        # real-valued statistics go through the same callback (optimizeP's pair sums)
        r = np.array([1.5 + rank, 2.0]); reduce_(0, None, r)
        assert r.tolist() == [4.0, 4.0]
        q.put((rank, "ok"))
    except Exception:   # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_hyper_statistics_reducer_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_hyper_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
