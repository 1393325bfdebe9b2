"""CPU tests of the drop-in boundary: libmvtm.so loads, exports every symbol include/mvtm.h declares, and fails
loudly (no fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mvtm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvtm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(engine_lib):
    from mvtopicmodel_b200 import _lib
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(engine_lib, n), f"{n} declared in mvtm.h but not exported by libmvtm.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_library_is_sm100a_native(engine_lib):
    from mvtopicmodel_b200 import _lib
    assert b"sm_100a" in engine_lib.mvtm_build_info()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass                 # TMA 1-D bulk copies of the n_wk rows
    assert "SYNCS.ARRIVE.TRANS64" in sass   # mbarrier expect_tx
    assert "REDG" in sass or "RED." in sass # count deltas as reductions
    # the DIRECT kernel (n_wk rows in registers): coherent 128-bit row loads and packed fp32 pairs
    direct = [b for b in sass.split("Function : ") if b.startswith("_Z19k_sweep_view_direct")]
    assert len(direct) >= 4                 # (1024,16) and (2048,32), single and multi view
    for b in direct:
        assert "LDG.E.128.STRONG.GPU" in b and "FFMA2" in b and "UBLKCP" not in b


def test_flag_constants_match_the_header():
    """The flag values the Python tests and tools pass as plain integers are the header's."""
    src = open(os.path.join(ROOT, "include", "mvtm.h")).read()
    flags = {k: int(v) for k, v in re.findall(r"#define\s+(MVTM_FLAG_[A-Z0-9_]+)\s+(\d+)u", src)}
    assert flags["MVTM_FLAG_DOC_ORDER"] == 1 and flags["MVTM_FLAG_SINGLE_WARP"] == 2
    assert flags["MVTM_FLAG_Q1_COMPAT"] == 4 and flags["MVTM_FLAG_BETA_MALLET"] == 8 and flags["MVTM_FLAG_TMA_RING"] == 16
    assert len(set(flags.values())) == len(flags) and all(v & (v - 1) == 0 for v in flags.values())


def test_no_cpu_fallback_without_device(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from mvtopicmodel_b200 import Engine, MvtmError
    import numpy as np
    with pytest.raises(MvtmError) as ei:
        Engine(10, [5], [(np.array([0, 2], dtype=np.int64), np.array([1, 2], dtype=np.int32))])
    assert ei.value.status == 3 and "no CPU fallback" in str(ei.value)


def test_argument_validation_before_any_device_work(engine_lib):
    from mvtopicmodel_b200 import _lib
    h = C.c_void_p()
    V = (C.c_int32 * 1)(5)
    cfg = _lib.MvtmConfig(0, 1, 1, V, 1, 0, 0, 0, 1, 0, 0, 0)          # K = 0
    assert engine_lib.mvtm_create(C.byref(cfg), C.byref(h)) == 1
    cfg = _lib.MvtmConfig(5000, 1, 1, V, 1, 0, 0, 0, 1, 0, 0, 0)       # K beyond this build
    assert engine_lib.mvtm_create(C.byref(cfg), C.byref(h)) == 5
    assert b"2048" in engine_lib.mvtm_last_error(None)
    assert engine_lib.mvtm_create(None, C.byref(h)) == 1
    assert engine_lib.mvtm_destroy(None) == 0


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under mvtopicmodel_b200/ or include/ may import, include or link it."""
    bad = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s*[\"<][^\">]*oracle)|(libmvtm_oracle)|(orc_[a-z_]+\s*\()", re.M)
    for top in ("mvtopicmodel_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(txt), f"{os.path.join(dirpath, f)} references the oracle"


def test_cpp_host_mirror_compiles(engine_lib, tmp_path):
    exe = tmp_path / "host_driver"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_driver.cpp"),
                           "-o", str(exe), "-L" + os.path.join(ROOT, "mvtopicmodel_b200"), "-lmvtm",
                           "-Wl,-rpath," + os.path.join(ROOT, "mvtopicmodel_b200")])
    import torch
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stdout     # loud failure, not a fallback


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the engine) runs without a GPU and prints ONE JSON line
    with the contract's keys; the sample is cut down so the test takes seconds."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample-docs", "1500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gibbs_token_updates_per_sec" and d["unit"] == "tokens/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


def test_launch_shapes_fit_the_sm(engine_lib):
    """Every launch shape the library would choose -- each K, 1..8 views, default / ring / compat kernels -- fits an SM of a B200:
    dynamic + static shared memory within 227 KB, and the warps of a CTA times the registers the chosen kernel was COMPILED with
    (cuobjdump -res-usage of the in-tree library) within the 16 K registers of each of the four sub-partitions.  (A 13-warp
    152-register shape fits the SM's 64 K registers on paper and still cannot launch: found on the GPU, now a CPU test.)"""
    from mvtopicmodel_b200 import _lib
    res = subprocess.run(["cuobjdump", "-res-usage", _lib.SO_PATH], capture_output=True, text=True).stdout
    regs = {}
    for name, r, shared in re.findall(r"Function (_Z\d+k_sweep_view\w+):\s*\n\s*REG:(\d+) STACK:\d+ SHARED:(\d+)", res):
        m = re.match(r"_Z\d+k_sweep_view(_direct)?ILi(\d+)ELi(\d+)ELb(\d)(?:ELb(\d))?", name)
        direct, KS, G, multi, q1 = bool(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4) == "1", m.group(5) == "1"
        regs[(direct, KS, G, multi, q1)] = (int(r), int(shared))
    assert len(regs) >= 48
    FLAG_Q1, FLAG_RING = 4, 16
    seen_direct = set()
    for K in list(range(1, 2049, 7)) + [128, 129, 256, 257, 384, 385, 500, 512, 513, 768, 769, 1000, 1024, 1025, 1536, 1537, 2000, 2048]:
        KS = next(128 * j for j in (1, 2, 3, 4, 6, 8, 12, 16) if 128 * j >= K)
        for M in (1, 2, 3, 4, 8):
            for flags in (0, FLAG_RING, FLAG_Q1):
                g, ring, warps, nreg = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
                smem = C.c_int64()
                rc = engine_lib.mvtm_test_launch_shape(K, M, flags, C.byref(g), C.byref(ring), C.byref(warps), C.byref(smem), C.byref(nreg))
                assert rc == 0, (K, M, flags, rc)
                direct = ring.value == 0
                assert not (direct and flags), (K, M, flags)                       # ring / compat requests never get the DIRECT kernel
                key = (direct, KS, g.value, M > 1, bool(flags & FLAG_Q1))
                assert key in regs, f"no compiled kernel for {key} (K={K}, M={M}, flags={flags})"
                r, static = regs[key]
                if direct:
                    assert r <= nreg.value
                    seen_direct.add((KS, g.value))
                assert 1 <= warps.value <= 32
                assert smem.value + static <= 227 * 1024, (K, M, flags, smem.value)
                per_partition = -(-warps.value // 4)                                # warps of the CTA on the fullest sub-partition
                assert per_partition * 32 * (-(-r // 8) * 8) <= 16384, (K, M, flags, warps.value, r)
    assert seen_direct == {(1024, 16), (2048, 32)}                                  # the shapes DIRECT is the default for
    assert engine_lib.mvtm_test_launch_shape(2049, 1, 0, C.byref(g), C.byref(ring), C.byref(warps), C.byref(smem), C.byref(nreg)) == 1
