"""Build libnsem_cuda.so (sm_100a) in-tree with nvcc.  `python -m nebulasem_b200.build [--force]`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libnsem_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, "nsem_cuda.cu")]
    deps = srcs + [os.path.join(CSRC, "nsem_kernels.cuh"), os.path.join(ROOT, "include", "nsem_c.h")]
    if not force and _newer(LIB, deps):
        return LIB
    cmd = [NVCC, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-DNSEM_WITH_NCCL",
           *srcs, "-o", LIB, "-lnccl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    print("[nebulasem_b200.build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
