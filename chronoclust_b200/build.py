"""Builds chronoclust_b200/libchronoclust_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "api.cu")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "common.cuh", "nearest.cuh", "online.cuh", "offline.cuh", "engine.cuh")]
DEPS.append(os.path.join(os.path.dirname(HERE), "include", "chronoclust_b200.h"))
SO = os.path.join(HERE, "libchronoclust_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",            # parity: the reference never contracts mul+add
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if os.environ.get("CCB_PTXAS_V") else "-O3",
    "--split-compile", "0",   # optimise the kernels of this one translation unit on all host cores
]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return SO
    cmd = [nvcc()] + NVCC_FLAGS + ["-o", SO, SRC]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
