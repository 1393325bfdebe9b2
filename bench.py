#!/usr/bin/env python
"""bench.py -- Gibbs token-updates/s of the B200 engine on the BASELINE.json target shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--docs D]

Default workload: acm_2v (BASELINE configs[2]: two views, K = 1000) at 1 000 000 documents IN TOTAL -- north_star's target
("a 1M-document, 2-view, 1000-topic synthetic corpus").  The corpus is the union of 8 blocks of D/8 documents; rank r of N
holds blocks r, r+N, ... so every N in {1, 2, 4, 8} samples the SAME corpus (strong scaling).  A "step" is one full Gibbs
sweep (every view) over the whole corpus.  N > 1 is launched by torchrun, one rank per GPU; n_wk is replicated and the
per-sweep count deltas are all-reduced with NCCL inside the timed region (overlapped with the other view's pass).  Rank 0
prints ONE JSON line.  Other workloads (lda_100k, pubmed_3v, stress_4v, uniform_k1000) via --workload; --docs = total documents.

--impl reference times the reference's CPU scheme (oracle/: multithreaded restatement of the Java worker/updater/queue
design; the Java code itself cannot run, there is no JVM in this image) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gibbs_token_updates_per_sec"
UNIT = "tokens/s"
N_BLOCKS = 8
DEFAULT_DOCS = {"acm_2v": 1_000_000, "lda_100k": 100_000, "pubmed_3v": 1_000_000, "stress_4v": 2_000_000, "uniform_k1000": 200_000}
CFG_IDX = {"lda_100k": 1, "acm_2v": 2, "pubmed_3v": 3, "stress_4v": 4}


def b_tok_view(K, m, mean_lens):
    """Algorithmic bytes per token update of view m, SURVEY.md 8(d) / BASELINE.md section 3: 4K (n_wk row) + 4 (word) + 8 (z r/w)
    + 16 (two count RMWs) + 8K/N_m (n_k and alpha rows per doc-view) + 4*sum_{i != m} N_i/N_m (other views' z re-read)."""
    n = mean_lens[m]
    other = sum(mean_lens[i] for i in range(len(mean_lens)) if i != m)
    return 4 * K + 28 + 8.0 * K / n + 4.0 * other / n


def b_tok(K, mean_lens, tokens):
    tot = float(sum(tokens))
    return sum(t * b_tok_view(K, m, mean_lens) for m, t in enumerate(tokens) if t) / tot


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v == "Active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---- corpus: 8 blocks, generated in parallel before CUDA is touched ---------------------------------------------------------
def _gen_block(args):
    workload, block, docs = args
    from mvtopicmodel_b200 import corpus
    if workload == "uniform_k1000":
        # the genuinely HBM-bound case: uniform words over V = 400 K at K = 1000 (1.6 GB table, no cache reuse)
        return corpus.generate_uniform(docs, 1000, 400_000, 100, seed=7 + block)
    return corpus.generate(workload, shard=block, docs=docs)


def build_corpus(workload, total_docs, rank, world):
    """This rank's share of the corpus: blocks rank, rank+world, ... of N_BLOCKS, concatenated (doc-aligned across views)."""
    per = total_docs // N_BLOCKS
    blocks = list(range(rank, N_BLOCKS, world)) if world <= N_BLOCKS else [rank % N_BLOCKS]
    jobs = [(workload, b, per) for b in blocks]
    if len(jobs) == 1:
        res = [_gen_block(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
            res = pool.map(_gen_block, jobs)
    K, Vs = res[0][0], res[0][1]
    views = []
    for m in range(len(Vs)):
        offs, words, base = [np.zeros(1, dtype=np.int64)], [], 0
        for r in res:
            off, w = r[2][m]
            offs.append(off[1:] + base); words.append(w); base += int(off[-1])
        views.append((np.concatenate(offs), np.concatenate(words)))
    return K, Vs, views


def cpu_reference_run(workload, sample_docs, steps, warmup, threads):
    """The reference's multithreaded CPU scheme (oracle restatement) on a bounded sample of the workload."""
    from oracle import oracle as O
    O.build()
    K, Vs, views = _gen_block((workload, 0, sample_docs))
    o = O.Oracle(K, Vs, views, seed=2026)
    o.init_assignments()
    o.rebuild_trees()
    ntok = sum(o.ntok)
    # the reference's own algorithm: F+trees maintained per delta by the updater threads (stale otherwise, Q3) and the dead
    # dense-index insertion (Q1) -- the configuration tests/test_reference_vectors.py shows to reproduce the reference's
    # sampler bytecode token for token
    flags = O.F_STALE_TREES | O.F_Q1_COMPAT
    for it in range(1, warmup + 1):
        o.sweep_mt(it, threads, flags)
    t0 = time.perf_counter()
    for it in range(warmup + 1, warmup + steps + 1):
        o.sweep_mt(it, threads, flags)
    dt = time.perf_counter() - t0
    assert o.check_invariants() == 0
    return ntok * steps / dt, dt / steps * 1e3, ntok, K


def load_traffic(workload, docs_per_gpu):
    """ncu evidence for the dominant kernel of this workload (profiles/ncu_traffic.json: one entry per captured workload)."""
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        allw = json.load(open(tp))
    except Exception:
        return None
    ent = allw.get(workload) if isinstance(allw, dict) else None
    if not isinstance(ent, dict):
        return None
    return ent


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--workload", default="acm_2v")
    ap.add_argument("--docs", type=int, default=None, help="documents IN TOTAL over all ranks (default: 1 M for acm_2v)")
    ap.add_argument("--cpu-sample-docs", type=int, default=25000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="N > 1: exchange the counts after the sweep instead of under it")
    ap.add_argument("--narrow-all", action="store_true", help="N > 1 with overlap: every view's exchange on the CTA-limited communicator")
    ap.add_argument("--reserve-sms", type=int, default=int(os.environ.get("MVTM_RESERVE_SMS", "8")),
                    help="N > 1 with overlap: SMs the persistent sweep kernel leaves to the collective")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from mvtopicmodel_b200 import corpus
    total_docs = args.docs if args.docs else DEFAULT_DOCS.get(args.workload, 100_000)
    total_docs = max(N_BLOCKS, total_docs // N_BLOCKS * N_BLOCKS)
    if args.workload == "uniform_k1000":
        shape = "uniform words, V=400000, K=1000, 100 tokens/doc (HBM-bound control: no cache reuse)"
    else:
        cfg = corpus.CONFIGS[args.workload]
        shape = (f"BASELINE configs[{CFG_IDX.get(args.workload, -1)}] synthetic shape, {len(cfg['views'])} view(s), "
                 f"V={[v[0] for v in cfg['views']]}, K={cfg['K']}")
    wl_desc = f"{args.workload}: {shape}; {total_docs} docs in total, sharded over the ranks"
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 5)); warmup = max(1, min(args.warmup, 1))
        val, ms, ntok, K = cpu_reference_run(args.workload, args.cpu_sample_docs, steps, warmup, threads)
        nst, nut = 3 * threads // 4, threads // 4
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32+f64",
                "data": "synthetic", "config": {"workload": wl_desc, "sample": f"{args.cpu_sample_docs} docs / {ntok} tokens per step"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"first {args.cpu_sample_docs} docs ({ntok} tokens) of {args.workload}, {steps} sweeps; "
                                           f"{nst} sampler + {nut} updater threads (M:1036-1037); C restatement of the Java scheme (token-for-token equal to the reference's sampler bytecode in its sequential form), no JVM in this image"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # corpus first: the block generators fork, which must happen before CUDA is initialised
    K, Vs, views = build_corpus(args.workload, total_docs, rank, world)
    M = len(views)
    D_local = len(views[0][0]) - 1
    # a single view has nothing to hide its exchange under (its next pass needs the result at once): serial there
    overlap = world > 1 and M > 1 and not args.no_overlap
    import torch
    from mvtopicmodel_b200 import Engine
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # torch.distributed is used for two things only: handing NCCL's 128-byte unique id to the ranks and the max-over-ranks of
        # the timings.  The count exchange itself runs inside libmvtm.so (mvtm_comm_init / mvtm_sweep_dist, NCCL behind the C ABI).
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    eng = Engine(K, Vs, views, seed=2026, device=local_rank, doc_id_base=rank, doc_id_stride=world,
                 max_ctas=(n_sms - args.reserve_sms) if overlap else 0)
    ntok_local = sum(eng.ntok)
    if world > 1:
        box = [Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world, hidden_ctas=args.reserve_sms if overlap else 0)
    eng.init_assignments()
    if world > 1:
        eng.sync_counts()               # local counts -> global counts, snapshots for the sum-form exchange

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def step(it):
        if world > 1:
            eng.sweep_dist(it)          # passes + exchange; view m's all-reduce runs under the following passes / the next sweep
        else:
            eng.sweep(it)

    def drain():
        if world > 1:
            eng.comm_drain()            # the last views' exchange belongs to the timed region

    it = 0
    for _ in range(6):          # set-up: the engine's ring-depth autotune takes three samples per depth (not warm-up, not timed)
        it += 1
        step(it)
    for _ in range(args.warmup):
        it += 1
        step(it)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    view_ms, changed = np.zeros(M), 0
    ev0.record()
    for _ in range(args.steps):
        it += 1
        step(it)
        st = eng.stats()
        view_ms += np.array(st["ms_view"])
        changed += st["changed"]
    drain()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    ring = eng.stats()["ring_locked"]
    if world > 1:
        ci = eng.comm_info()
        allreduce_bytes, nccl_version = ci["bytes_last_sweep"], ci["nccl_version"]
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([ms_total] + list(view_ms), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, view_ms_max = float(t[0]), t[1:].cpu().numpy()
        n = torch.tensor([ntok_local], device="cuda", dtype=torch.int64)
        dist.all_reduce(n)
        ntok_global = int(n[0])
    else:
        ntok_global, view_ms_max = ntok_local, view_ms
    if world == 1:
        viol = eng.check_invariants()
    else:
        # replicas hold GLOBAL counts: per view, sum_t n_k must equal the token total over all ranks and be identical everywhere
        viol = 0
        for m in range(M):
            nk = torch.from_numpy(eng.get_counts(m, want_nwk=False)[1].astype(np.int64)).cuda()
            tot = torch.tensor([eng.ntok[m]], device="cuda", dtype=torch.int64)
            dist.all_reduce(tot)
            mx, mn = nk.clone(), nk.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            viol += int(nk.sum().item() != tot.item()) + int((mx != mn).sum().item()) + int((nk < 0).sum().item())
    value = ntok_global * args.steps / (ms_total * 1e-3)

    # end-to-end: the same sweep through HOST buffers (pinned), H2D + count rebuild + sweep + D2H inside the timing
    e2e = None
    if not args.no_e2e:
        zh = [torch.empty(n, dtype=torch.int32).pin_memory() for n in eng.ntok]
        zn = [z.numpy() for z in zh]
        for m in range(M):
            zn[m][:] = eng.get_assignments(m)

        def e2e_step(i):
            if world > 1:
                eng.sweep_host_dist(i, zn)                 # upload + compare (or recount + all-reduce), passes + exchanges, z written back
            else:
                eng.sweep_host(i, zn)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            it += 1
            e2e_step(it)
        barrier()
        ev0.record()
        for _ in range(e2e_steps):
            it += 1
            e2e_step(it)
        ev1.record()
        barrier()
        e2e_ms = ev0.elapsed_time(ev1)
        if dist:
            t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t[0])
        e2e = {"value": ntok_global * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * ntok_local,
               "d2h_bytes_per_step": 4 * ntok_local, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
               "call": "mvtm_sweep_host (pinned host z in/out: chunked upload + count rebuild, sweep, new z stored to the pinned arrays by the kernel)" if world == 1 else
                       "mvtm_sweep_host_dist (per rank: pinned host z in/out; the upload is compared with the resident assignments, a 4-byte all-reduce tells every rank "
                       "whether all shards came back unchanged -- they do here -- so the resident global counts are kept (any difference: recount + one all-reduce per view); "
                       "passes with overlapped exchanges; new z stored to the pinned arrays by the kernel)"}
        for m in range(M):
            assert np.array_equal(zn[m], eng.get_assignments(m)), "host arrays differ from the device assignments"

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    mean_lens = [eng.ntok[m] / max(1, int(((views[m][0][1:] - views[m][0][:-1]) > 0).sum())) for m in range(M)]
    btok_all = b_tok(K, mean_lens, eng.ntok)
    # the dominant kernel: the view pass with the most device time (k_sweep_view of that view; one launch per step)
    dom = int(np.argmax(view_ms_max))
    btok_dom = b_tok_view(K, dom, mean_lens)
    launch_ms = float(view_ms_max[dom]) / args.steps
    achieved = btok_dom * eng.ntok[dom] / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else 0.0
    ent = load_traffic(args.workload, D_local)
    traffic = l2_hit = dram_frac = None
    bound = "unknown (no ncu capture of this workload under profiles/)"
    if ent:
        # the capture's DRAM bytes are per launch at ITS document count: scale per token to this launch
        per_tok = ent["dram_bytes_per_launch"] / ent["tokens_per_launch"]
        traffic = per_tok * eng.ntok[dom]
        l2_hit = ent.get("l2_hit_rate_pct")
        dram_frac = traffic / (launch_ms * 1e-3) / 1e9 / peak if launch_ms > 0 else None
        bound = ent.get("bound", "hbm" if dram_frac and dram_frac > 0.6 else "l2+issue")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32+f32", "data": "synthetic",
            "config": {"workload": wl_desc, "docs_total": total_docs, "docs_per_gpu": D_local, "tokens_per_gpu": ntok_local,
                       "tokens_total": ntok_global, "K": K, "views": M, "ring_depth": ring,
                       "l2": "inputs larger than L2 (z+words+n_wk = %d MB per sweep vs 126 MB L2), no explicit flush" %
                             int((8 * ntok_local + 4 * sum(Vs) * eng.row_stride()) / 1e6),
                       "changed_frac": changed / max(1, ntok_local * args.steps), "invariant_violations": viol},
            # `achieved` counts ALGORITHMIC bytes (one dense n_wk row per token); `bound` says what ncu shows limits the kernel:
            # on Zipf corpora the hot rows are served from L2 and DRAM moves only `traffic` bytes per launch (dram_frac of peak)
            "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "dram_frac": dram_frac, "l2_hit": l2_hit, "traffic": traffic,
                         "kernel": ("k_sweep_view_direct (n_wk rows in registers)" if ring[dom] == 0 else "k_sweep_view (TMA ring)") + f", view {dom}", "bytes_per_token": btok_dom, "tokens_per_launch": eng.ntok[dom],
                         "avg_launch_ms": launch_ms, "bytes_per_token_all_views": btok_all,
                         "whole_job_frac": btok_all * value / 1e9 / (peak * world),
                         "peak_source": peak_src, "traffic_source": ent.get("source") if ent else None},
            "e2e": e2e, "clocks": clocks,
                    # my kernels inside the timed region: one k_sweep_view per view and step, plus the exchange's finishing pass
            # (one fused k_finish_sum_exchange4 per view)
            "gpu_launches": args.steps * M * (1 + (1 if world > 1 else 0))}
    if world > 1:
        line["config"]["allreduce_bytes_per_sweep"] = allreduce_bytes
        line["config"]["exchange"] = ("inside libmvtm.so (mvtm_sweep_dist, NCCL %s): " % nccl_version) + (
            f"overlapped -- view m's all-reduce under the following passes, sweep grid {n_sms - args.reserve_sms} CTAs, hidden exchanges on a "
            f"communicator with maxCTAs={args.reserve_sms}, the longest view's on the wide one" if overlap else "after the pass (single view: nothing to hide it under)")
    if world == 1 and not args.no_cpu_baseline:
        val, ms, ntok_s, _ = cpu_reference_run(args.workload, args.cpu_sample_docs, 3, 1, threads)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"first {args.cpu_sample_docs} docs ({ntok_s} tokens) of {args.workload}, 3 sweeps, "
                                          f"{3 * threads // 4} sampler + {threads // 4} updater threads"}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
