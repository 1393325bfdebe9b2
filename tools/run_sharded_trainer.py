"""Multi-GPU estimate() through dist.ShardedTrainer over NCCL (launch with torchrun, one rank per GPU):
overlapped count exchange (hidden views on a CTA-limited communicator), hyper-parameter step on all-reduced statistics,
global log-likelihood.  Prints the LL/token series and checks that every rank ends with the same hyper-parameters and counts.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29620 tools/run_sharded_trainer.py
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from mvtopicmodel_b200 import corpus
    from mvtopicmodel_b200.dist import ShardedTrainer
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    workload, docs = (sys.argv[1] if len(sys.argv) > 1 else "pubmed_3v"), int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    K, Vs, views = corpus.generate(workload, shard=rank, docs=docs)
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = 8
    opts.config.min_ctas = 1
    narrow = dist.new_group(pg_options=opts)
    t = ShardedTrainer(K, Vs, views, rank, world, device=local, seed=2026, narrow_group=narrow, overlap=True,
                       stage_device=torch.device(f"cuda:{local}"))
    ntok = torch.tensor([float(n) for n in t.engine.ntok], device="cuda", dtype=torch.float64)
    dist.all_reduce(ntok)
    ntok = ntok.cpu().numpy()
    t0 = time.perf_counter()
    t.estimate(iters, burninPeriod=10, optimizeInterval=10, ll_every=10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    hf = t.engine.get_hyper_full()
    dig = hashlib.sha256()
    for k in ("alpha", "alphaSum", "beta", "betaSum", "gamma", "p_a"):
        dig.update(np.ascontiguousarray(hf[k]).tobytes())
    for m in range(len(Vs)):
        dig.update(t.engine.get_counts(m, want_nwk=False)[1].tobytes())
    digs = [None] * world
    dist.all_gather_object(digs, dig.hexdigest())
    tot = [int(t.engine.get_counts(m, want_nwk=False)[1].sum()) for m in range(len(Vs))]
    if rank == 0:
        print("LL/token:", [(it, (ll / ntok).round(4).tolist()) for it, ll in t.ll_series])
        print("ranks agree on hyper-parameters and counts:", len(set(digs)) == 1, "| n_k totals", tot, "== tokens", [int(x) for x in ntok],
              "| inactive topics", len(hf["inactive"]), "| gamma", np.round(hf["gamma"], 4).tolist(), "beta", np.round(hf["beta"], 5).tolist())
        print("wall %.2f s for %d iterations incl. %d hyper-parameter steps and %d LL evaluations" % (dt, iters, max(0, (iters - 10) // 10), iters // 10))
        assert len(set(digs)) == 1 and tot == [int(x) for x in ntok]
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
