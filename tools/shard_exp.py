"""Experiment: LL/token of the sharded trainer with 1 and 2 ranks sharing cuda:0 (gloo), with and without the hyper-parameter step."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def worker(rank, world, port, q, opt):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvtopicmodel_b200 import corpus
    from mvtopicmodel_b200.dist import ShardedTrainer
    big = os.environ.get("SHARD_EXP_BIG") == "1"
    K, Vs, full = corpus.generate(dict(D=40_000, K=100, views=[(5000, 40, 0.6, 1.0, 512), (800, 6, 0.5, 0.8, 64)])) if big else corpus.generate("small_3v")
    t = ShardedTrainer(K, Vs, corpus.shard_views(full, rank, world), rank, world, device=0, seed=31, max_ctas=(64 if big else 8) // world, warps_per_cta=4)
    ll0 = t.global_loglik()
    t.estimate(60 if big else 200, burninPeriod=20, optimizeInterval=opt, ll_every=20)
    t.ll_series.insert(0, (0, ll0))
    # exact recount of the global tables from every rank's assignments
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import recount
    zs_all = [None] * world
    dist.all_gather_object(zs_all, [t.engine.get_assignments(m) for m in range(len(Vs))])
    bad = 0
    for m in range(len(Vs)):
        nwk = np.zeros((Vs[m], K), dtype=np.int64); nk = np.zeros(K, dtype=np.int64)
        for r in range(world):
            vr = corpus.shard_views(full, r, world)
            (a, b), = recount([vr[m]], [zs_all[r][m]], K, [Vs[m]])
            nwk += a; nk += b
        a, b = t.engine.get_counts(m)
        bad += int((a != nwk).sum()) + int((b != nk).sum())
    t.ll_series.append((-1, np.array([bad] * len(Vs), dtype=np.float64) * np.array([len(v[1]) for v in full])))
    ntok = np.array([len(v[1]) for v in full], dtype=np.float64)
    if rank == 0:
        q.put([(it, (ll / ntok).round(5 if it == 0 else 3).tolist()) for it, ll in t.ll_series])
    dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = 29700
    for opt in ((0,) if os.environ.get('SHARD_EXP_BIG') == '1' else (0, 20)):
        for world in (1, 2):
            q = ctx.Queue(); port += 1
            ps = [ctx.Process(target=worker, args=(r, world, port, q, opt)) for r in range(world)]
            [p.start() for p in ps]
            print("opt", opt, "world", world, q.get(timeout=900), flush=True)
            [p.join() for p in ps]
