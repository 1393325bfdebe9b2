"""Build libmvtm.so (the C-ABI CUDA engine) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "mvtm.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "mvtm_kernels.cuh"), os.path.join(HERE, "csrc", "mvtm_optim.inl"),
        os.path.join(HERE, "csrc", "mvtm_comm.inl"),
        os.path.join(ROOT, "include", "mvtm.h")]
OUT = os.path.join(HERE, "libmvtm.so")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-I" + os.path.join(ROOT, "include"), "-shared", "-Xcompiler", "-fPIC", "-ldl"]
# NOT -split-compile: with it ptxas' code for the sweep kernel changes from build to build of the SAME source (three variants were
# seen, one 27 % slower: profiles/r2_ab_codegen_variants.log); the single-threaded compile is deterministic (~110 s).


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
