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
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "nsem_c.h")]
    if not force and _newer(LIB, deps):
        return LIB
    cmd = [NVCC, "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared", "-DNSEM_WITH_NCCL",
           *srcs, "-o", LIB, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    print("[nebulasem_b200.build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


HOST_LIB = os.path.join(LIBDIR, "libnsem_host.so")
EULER_BIN = os.path.join(LIBDIR, "euler")
CXX = os.environ.get("NSEM_CXX", "g++")   # the image's $CXX wrapper (/opt/gcc) lacks libgomp.spec
# -ffp-contract=off: geometry and reference-state arithmetic must round like the reference's own -O2 build
HOST_FLAGS = ["-O3", "-std=c++17", "-ffp-contract=off", "-fPIC", "-fopenmp", "-Wall", "-Wno-unknown-pragmas"]


def build_host(force: bool = False) -> str:
    hdir = os.path.join(CSRC, "host")
    srcs = [os.path.join(hdir, f) for f in ("io.cpp", "mesh.cpp", "dg.cpp", "partition.cpp", "euler_app.cpp", "amr.cpp", "capi_host.cpp")]
    metis = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "targets", "x86_64-linux", "lib", "libmetis_static.a")
    metis_flags = ["-DNSEM_WITH_METIS"] if os.path.exists(metis) else []
    metis_libs = [metis] if os.path.exists(metis) else []
    deps = srcs + [os.path.join(hdir, "nsem_host.h"), os.path.join(ROOT, "include", "nsem_c.h"), LIB]
    if force or not _newer(HOST_LIB, deps):
        cmd = [CXX, *HOST_FLAGS, *metis_flags, "-shared", *srcs, *metis_libs, "-o", HOST_LIB, "-L" + LIBDIR, "-lnsem_cuda",
               "-Wl,-rpath,$ORIGIN"]
        print("[nebulasem_b200.build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    main = os.path.join(hdir, "euler_main.cpp")
    if force or not _newer(EULER_BIN, [main, HOST_LIB]):
        cmd = [CXX, *HOST_FLAGS, main, "-o", EULER_BIN, "-L" + LIBDIR, "-lnsem_host", "-lnsem_cuda", "-Wl,-rpath,$ORIGIN"]
        print("[nebulasem_b200.build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    # the same program under the reference's other app name on this path (apps/convection): the solver is chosen by the controls file
    conv = os.path.join(LIBDIR, "convection")
    if force or not _newer(conv, [EULER_BIN]):
        import shutil
        shutil.copy2(EULER_BIN, conv)
    return HOST_LIB


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
