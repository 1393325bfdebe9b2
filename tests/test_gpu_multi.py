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
@pytest.mark.parametrize("world", [2, 4, 8])
def test_cpp_dist_driver_without_torch(engine_lib, tmp_path, world):
    """tests/cpp/dist_driver.cpp: one host thread and one handle per GPU, no torch and no Python in the process.  Checks that every
    rank's replica equals the histogram of ALL ranks' assignments after overlapped mvtm_sweep_dist sweeps, after a stateless
    mvtm_sweep_host_dist step + mvtm_sync_counts(rebuild), and after a hyper-parameter step whose statistics were reduced inside
    the library; the global log-likelihood is bit-identical on every rank and improved."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    exe = tmp_path / "dist_driver"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dist_driver.cpp"),
                           "-o", str(exe), "-L" + os.path.join(ROOT, "mvtopicmodel_b200"), "-lmvtm", "-lpthread",
                           "-Wl,-rpath," + os.path.join(ROOT, "mvtopicmodel_b200")])
    r = subprocess.run([str(exe), str(world), "8"], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0 and "OK" in r.stdout
