"""Multi-GPU host logic: documents sharded over ranks, n_wk / n_k replicated, one integer all-reduce of the
per-sweep count deltas per view.  torch.distributed is plumbing only (NCCL on GPUs,
gloo in the CPU tests); the delta arithmetic runs in the engine's own kernels (mvtm_delta_*).

Protocol per view, bit-exact in int32:
    delta_g = replica_g - snapshot          (mvtm_delta_export, in place)
    delta   = all_reduce_sum(delta_g)       (NCCL over NVLink / NVSwitch)
    replica = snapshot + delta; snapshot = replica   (mvtm_delta_import)
so after the exchange every rank holds the same global counts, equal to the histogram of all ranks' assignments.
"""
import numpy as np


def shard_doc_ids(num_docs, rank, world):
    """Strided sharding: rank r owns global documents r, r+world, ... (all views of a document stay together).
    Returns (doc_id_base, doc_id_stride, local_count) -- the engine keys its Philox stream on the global id."""
    return rank, world, len(range(rank, num_docs, world))


def make_stat_reducer(group=None, device=None):
    """The callback mvtm_set_stat_reducer expects, over torch.distributed: all-reduce (sum / max) of the optimiser's host-side
    statistics, in place.  `device`: stage through that CUDA device (NCCL groups cannot reduce host tensors); None for gloo."""
    def fn(op, ints, reals):
        import torch
        import torch.distributed as dist
        rop = dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX
        for arr in (ints, reals):
            if arr is None:
                continue
            t = torch.from_numpy(arr)
            if device is None:
                dist.all_reduce(t, op=rop, group=group)
            else:
                g = t.to(device)
                dist.all_reduce(g, op=rop, group=group)
                t.copy_(g.cpu())
    return fn


class _DevBuf:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


class EngineAdapter:
    """delta_begin / delta_reset / delta_export(m) -> (n_wk tensor, n_k tensor) / delta_import(m) on a CUDA Engine."""

    def __init__(self, engine, device):
        import torch
        self.e, self.dev, self.torch = engine, device, torch
        self.M = engine.M

    def delta_begin(self):
        self.e.delta_begin()

    def delta_reset(self):
        self.e.delta_reset()

    def delta_export(self, m):
        (p1, n1), (p2, n2) = self.e.delta_export(m)
        t = self.torch
        return (t.as_tensor(_DevBuf(p1, n1), device=f"cuda:{self.dev}"), t.as_tensor(_DevBuf(p2, n2), device=f"cuda:{self.dev}"))

    def delta_import(self, m):
        self.torch.cuda.synchronize(self.dev)      # the all-reduce ran on torch's stream
        self.e.delta_import(m)

    def sum_buffers(self, m):
        (p1, n1), (p2, n2) = self.e.sum_exchange_buffers(m)
        t = self.torch
        return (t.as_tensor(_DevBuf(p1, n1), device=f"cuda:{self.dev}"), t.as_tensor(_DevBuf(p2, n2), device=f"cuda:{self.dev}"))

    def sum_finish(self, m, world):
        self.torch.cuda.synchronize(self.dev)
        self.e.sum_exchange_finish(m, world)


class OverlapAdapter:
    """Engine side of OverlappedSweep: a side stream `comm` for the collective and its finishing pass, hand-over of each
    view's tables between the engine's stream and `comm` through the mvtm_*_wait_* entry points (device-side waits only)."""

    def __init__(self, engine, device, finish_ctas=0):
        import contextlib
        import torch
        self.e, self.dev, self.torch, self.M = engine, device, torch, engine.M
        self.comm = torch.cuda.Stream(device=device)
        self.finish_ctas = finish_ctas
        self._ctx = contextlib
        self.bufs = []
        for m in range(self.M):
            (p1, n1), (p2, n2) = engine.sum_exchange_buffers(m)
            assert p2 == p1 + 4 * n1, "n_k must follow n_wk in the same allocation"
            self.bufs.append(torch.as_tensor(_DevBuf(p1, n1 + n2), device=f"cuda:{device}"))

    def whole_buffer(self, m):
        return self.bufs[m]

    def sweep_view_async(self, it, m):
        self.e.sweep_view_async(it, m, 1)

    def comm_wait_view(self, m):
        self.e.stream_wait_view(m, self.comm.cuda_stream)

    def comm_context(self):
        return self.torch.cuda.stream(self.comm)

    def finish_async(self, m, world):
        self.e.sum_exchange_finish_async(m, world, self.comm.cuda_stream, self.finish_ctas)

    def view_wait_comm(self, m):
        self.e.view_wait_stream(m, self.comm.cuda_stream)

    def sweep_finish(self):
        self.e.sweep_finish()

    def drain(self):
        self.comm.synchronize()


class OverlappedSweep:
    """One Gibbs sweep over all views with the count exchange of view m running WHILE the following views (and the next
    sweep's earlier views) are sampled: view m's global counts are only needed when view m is sampled again (the reference's
    barrier M:1231 asks no more).  Sum-form protocol, one all-reduce per view (table and totals are one buffer).  Needs
    identical global counts on every rank at entry (i.e. after any completed exchange) and snapshots (delta_begin/reset)."""

    def __init__(self, adapter, group=None, view_groups=None):
        """view_groups: optional {view: process group} overriding `group` per view.  A view whose exchange cannot be hidden
        (the view with the longest pass: only the other, shorter passes lie between two of its own) should use a
        communicator that may take every SM, the hidden ones a communicator limited to the SMs the sweep leaves free."""
        import torch.distributed as dist
        self.a, self.dist, self.group = adapter, dist, group
        self.view_groups = dict(view_groups or {})
        self.bytes_per_exchange = 0

    def step(self, it):
        a, world = self.a, self.dist.get_world_size(self.group)
        total = 0
        for m in range(a.M):
            a.sweep_view_async(it, m)
            a.comm_wait_view(m)
            buf = a.whole_buffer(m)
            with a.comm_context():
                self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.view_groups.get(m, self.group))
            a.finish_async(m, world)
            a.view_wait_comm(m)
            total += buf.numel() * 4
        a.sweep_finish()
        self.bytes_per_exchange = total
        return total


class CountExchange:
    """Runs the delta protocol over a torch.distributed process group for any adapter with the four methods above."""

    def __init__(self, adapter, group=None):
        import torch.distributed as dist
        self.a, self.dist, self.group = adapter, dist, group
        self.bytes_per_exchange = 0

    def begin(self):
        self.a.delta_begin()

    def reset(self):
        self.a.delta_reset()

    def exchange_sum(self, views=None):
        """Sum-form exchange (one memory pass cheaper than `exchange`): valid whenever all ranks entered the sweep with the
        same global counts, i.e. after any completed exchange.  Adapter needs sum_buffers / sum_finish."""
        world = self.dist.get_world_size(self.group)
        total = 0
        for m in (range(self.a.M) if views is None else views):
            nwk, nk = self.a.sum_buffers(m)
            self.dist.all_reduce(nwk, op=self.dist.ReduceOp.SUM, group=self.group)
            self.dist.all_reduce(nk, op=self.dist.ReduceOp.SUM, group=self.group)
            total += (nwk.numel() + nk.numel()) * 4
            self.a.sum_finish(m, world)
        self.bytes_per_exchange = total
        return total

    def exchange(self, views=None):
        total = 0
        for m in (range(self.a.M) if views is None else views):
            nwk, nk = self.a.delta_export(m)
            self.dist.all_reduce(nwk, op=self.dist.ReduceOp.SUM, group=self.group)
            self.dist.all_reduce(nk, op=self.dist.ReduceOp.SUM, group=self.group)
            total += (nwk.numel() + nk.numel()) * 4
            self.a.delta_import(m)
        self.bytes_per_exchange = total
        return total


class ShardedTrainer:
    """One rank of a multi-GPU `estimate()` (M:1146-1239): this rank's shard of the documents (strided: rank, rank+world, ...),
    replicated count tables exchanged after every sweep (overlapped with sampling for multi-view corpora when `overlap`),
    the hyper-parameter step on statistics reduced over all ranks (mvtm_set_stat_reducer), the log-likelihood summed over ranks.
    Every rank ends each iteration with identical counts and hyper-parameters.

    `views` are the shard's views (corpus.shard_views); `group` / `narrow_group`: process groups for the exchanges (the narrow one
    is the CTA-limited communicator of the hidden views, see OverlappedSweep); `stage_device`: CUDA device to stage host
    statistics through when the group is NCCL (None for gloo)."""

    def __init__(self, K, Vs, views, rank, world, device=0, seed=1, group=None, narrow_group=None, overlap=False, reserve_sms=8,
                 stage_device=None, max_ctas=0, warps_per_cta=0):
        import torch
        import torch.distributed as dist
        from .engine import Engine
        self.K, self.M, self.rank, self.world, self.dist, self.group, self.torch = K, len(views), rank, world, dist, group, torch
        self.device = device
        n_sms = torch.cuda.get_device_properties(device).multi_processor_count
        self.overlap = bool(overlap and self.M > 1 and world > 1)
        if self.overlap:
            max_ctas = min(max_ctas, n_sms - reserve_sms) if max_ctas else n_sms - reserve_sms
        self.engine = e = Engine(K, Vs, views, seed=seed, device=device, doc_id_base=rank, doc_id_stride=world,
                                 max_ctas=max_ctas, warps_per_cta=warps_per_cta)
        e.set_stat_reducer(make_stat_reducer(group, stage_device) if world > 1 else None)
        self.xch = CountExchange(EngineAdapter(e, device), group)
        self.xch.reset()
        e.init_assignments()                       # M:465-515 on the global document ids: draws what the unsharded run draws
        self.xch.exchange()                        # local counts -> global counts
        self.ovl = None
        if self.overlap:
            critical = int(np.argmax(e.ntok))
            vg = {m: narrow_group for m in range(self.M) if m != critical} if narrow_group is not None else None
            self._ovl_adapter = OverlapAdapter(e, device)
            self.ovl = OverlappedSweep(self._ovl_adapter, group, view_groups=vg)
        self.burninPeriod, self.optimizeInterval = 200, 50
        self.ll_series = []

    def sweep(self, iteration):
        if self.ovl:
            self.engine.activate_topics()          # on the global counts of the previous sweep (waits for its exchanges; no-op
            self.ovl.step(iteration)               #   unless optimizeDP left inactive topics)
        else:
            self.engine.sweep(iteration)
            self.xch.exchange_sum()
            self.engine.activate_topics()          # U:263-270 on the global counts: every rank takes the same decision

    def drain(self):
        if self.ovl:
            self._ovl_adapter.drain()

    def global_loglik(self, quirk_len2=False):
        """modelLogLikelihood (M:3322-3452) of the whole corpus: document parts summed over ranks + the topic-word part."""
        self.drain()
        doc, word = self.engine.loglik_parts(quirk_len2)
        t = self.torch.from_numpy(doc.copy())
        dev = self.torch.device("cuda", self.device) if self.dist.get_backend(self.group) == "nccl" else None
        if dev is not None:
            g = t.to(dev); self.dist.all_reduce(g, group=self.group); t = g.cpu()
        else:
            self.dist.all_reduce(t, group=self.group)
        return t.numpy() + word

    def estimate(self, numIterations, burninPeriod=200, optimizeInterval=50, ll_every=10):
        e = self.engine
        M = self.M
        e.set_hyper(p_a=np.full((M, M), 0.2), p_b=np.ones((M, M)))                                    # M:1055-1058
        for iteration in range(1, numIterations + 1):
            if iteration < burninPeriod and M > 1:
                e.set_hyper(p_a=np.full((M, M), min(iteration / 100.0 + 0.3, 1.1)))                  # M:1166-1169
            elif iteration > burninPeriod and optimizeInterval != 0 and iteration % optimizeInterval == 0:
                self.drain()
                e.optimize_hyper(iteration)                                                          # M:1173-1210, global statistics
            self.sweep(iteration)
            if ll_every and iteration % ll_every == 0:
                self.ll_series.append((iteration, self.global_loglik()))
        self.drain()
