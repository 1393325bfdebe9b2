"""Multi-GPU through the C ABI alone (NCCL behind mvtm_comm_init): needs >= 2 visible GPUs, skipped otherwise (the driver's
round-end GPU pass runs on one GPU; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world,n_views", [(2, 2), (2, 1), (4, 2), (8, 2), (8, 1)])
def test_cpp_dist_driver_without_torch(engine_lib, tmp_path, world, n_views):
    """tests/cpp/dist_driver.cpp: one host thread and one handle per GPU, no torch and no Python in the process.  Checks that every
    rank's replica equals the histogram of ALL ranks' assignments after overlapped mvtm_sweep_dist sweeps, after mvtm_sweep_host_dist
    steps through host arrays (first call recounts, an unchanged call keeps the resident counts, one token edited on one rank makes
    every rank recount), and after a hyper-parameter step whose statistics were reduced inside
    the library; the global log-likelihood is bit-identical on every rank and improved.  n_views = 1: the serial exchange of a
    single-view corpus (nothing to hide it under)."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    exe = tmp_path / "dist_driver"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dist_driver.cpp"),
                           "-o", str(exe), "-L" + os.path.join(ROOT, "mvtopicmodel_b200"), "-lmvtm", "-lpthread",
                           "-Wl,-rpath," + os.path.join(ROOT, "mvtopicmodel_b200")])
    r = subprocess.run([str(exe), str(world), "8", str(n_views)], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0 and "OK" in r.stdout


def test_native_comm_layer_single_rank(engine_lib):
    """The library-side NCCL layer on ONE GPU (world = 1: every collective is the identity, every code path still runs): the
    communicator is created from a unique id, mvtm_sync_counts / mvtm_sweep_dist / mvtm_comm_drain keep the count invariants bit
    for bit and sample exactly what mvtm_sweep samples from the same state (same Philox keys, single-warp launch), the stateless
    mvtm_sweep_host_dist returns the device assignments -- recounting on its first call and whenever the caller's arrays or the
    handle's state changed in between, keeping the resident counts otherwise, with the same result either way -- and leaves global
    counts behind, mvtm_loglik_dist equals mvtm_loglik, mvtm_optimize_hyper runs through the library's own reducer."""
    import numpy as np
    from helpers import random_corpus
    from mvtopicmodel_b200 import Engine, MvtmError
    K, Vs = 130, [300, 100, 50]
    views = random_corpus(7, 600, K, Vs, [20, 4, 2], oov=True)
    a = Engine(K, Vs, views, seed=3, flags=2, ring_depth=1)              # MVTM_FLAG_SINGLE_WARP: deterministic
    b = Engine(K, Vs, views, seed=3, flags=2, ring_depth=1)
    uid = Engine.comm_unique_id()
    assert len(uid) == 128
    with pytest.raises(MvtmError):
        a.sweep_dist(1)                                                   # no communicator yet
    a.comm_init(uid, 0, 1, hidden_ctas=8)
    info = a.comm_info()
    assert info["rank"] == 0 and info["world"] == 1 and info["nccl_version"] >= 21800
    a.init_assignments(); b.init_assignments()
    with pytest.raises(MvtmError):
        a.sweep_dist(1)                                                   # counts not synchronised yet
    a.sync_counts()
    for it in range(1, 5):
        a.sweep_dist(it); b.sweep(it)
        a.comm_drain()
        assert a.check_invariants() == 0
        for m in range(3):
            assert np.array_equal(a.get_assignments(m), b.get_assignments(m))
    assert np.allclose(a.loglik_dist(), b.loglik(), rtol=1e-13, atol=0)
    z = [a.get_assignments(m).copy() for m in range(3)]
    a.sweep_host_dist(5, z); b.sweep(5)
    assert not a.last_host_step_used_resident_counts()                    # first host step: counts rebuilt from the upload
    for m in range(3):
        assert np.array_equal(z[m], a.get_assignments(m)) and np.array_equal(z[m], b.get_assignments(m))
    a.sweep_host_dist(6, z); b.sweep(6)                                   # arrays came back untouched: resident counts kept
    assert a.last_host_step_used_resident_counts()
    for m in range(3):
        assert np.array_equal(z[m], a.get_assignments(m)) and np.array_equal(z[m], b.get_assignments(m))
    # the caller edits its arrays: the step must notice, recount, and still sample what a resident engine fed the same edit samples
    z[0][:50] = (z[0][:50] + 1) % K
    for m in range(3):
        b.set_assignments(m, z[m])
    a.sweep_host_dist(7, z); b.sweep(7)
    assert not a.last_host_step_used_resident_counts()
    for m in range(3):
        assert np.array_equal(z[m], a.get_assignments(m)) and np.array_equal(z[m], b.get_assignments(m))
    # something else writes the handle's state in between (same assignments, so a comparison alone would not notice)
    for m in range(3):
        a.set_assignments(m, z[m])
    a.sweep_host_dist(8, z); b.sweep(8)
    assert not a.last_host_step_used_resident_counts()
    a.sweep_host_dist(9, z); b.sweep(9)
    assert a.last_host_step_used_resident_counts()
    for m in range(3):
        assert np.array_equal(z[m], a.get_assignments(m)) and np.array_equal(z[m], b.get_assignments(m))
    a.sweep_dist(10); b.sweep(10); a.comm_drain()                         # global counts were left behind: no mvtm_sync_counts needed
    assert a.check_invariants() == 0 and all(np.array_equal(a.get_assignments(m), b.get_assignments(m)) for m in range(3))
    a.optimize_hyper(60); b.optimize_hyper(60)
    ha, hb = a.get_hyper_full(), b.get_hyper_full()
    for k in ("alpha", "gamma", "beta", "p_a"):
        assert np.array_equal(ha[k], hb[k]), k
    a.close(); b.close()
