"""Links the PRODUCT's engine object (dorylus_b200/_obj/engine_cu.o, compiled by nvcc for sm_100a --
not recompiled, not modified) and the host-only objects (loader, partition) against a host-memory fake
of the CUDA runtime and scalar CPU statements of the kernel launchers:

    tests/hostcheck/_build/libdorylus_hostcheck.so

TEST INFRASTRUCTURE ONLY.  It exists so that the engine's host logic (tensor tables, operator order,
the apply-first schedule, window passes, error paths) can be exercised by `pytest -m "not gpu"` in a
container without a GPU.  The product binding (dorylus_b200/_lib.py) opens
dorylus_b200/libdorylus_b200.so by absolute path and knows nothing about this library.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OBJ = os.path.join(ROOT, "dorylus_b200", "_obj")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdorylus_hostcheck.so")
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build(force: bool = False) -> str:
    from dorylus_b200 import build as product_build

    product_build.build()  # makes sure engine_cu.o / loader_cpp.o / partition_cpp.o are current
    objs = [os.path.join(OBJ, n) for n in ("engine_cu.o", "loader_cpp.o", "partition_cpp.o", "tile_plan_cpp.o")]
    srcs = [os.path.join(HERE, n) for n in ("fake_cudart.cpp", "cpu_kernels.cpp")]
    hdrs = [os.path.join(ROOT, "dorylus_b200", "csrc", n) for n in ("common.cuh", "comm.h", "gat.cuh", "gemm_tc.cuh")]
    deps = objs + srcs + hdrs + [os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # -Bsymbolic: the engine's calls to cudaMalloc & co. must bind to the fakes in THIS library even when
    # a real libcudart is already in the process's global scope (torch loads its CUDA libraries RTLD_GLOBAL)
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-Wl,-Bsymbolic", "-I" + CUDA_INC, "-o", LIB] + srcs + objs + ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostcheck build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


def build_launchcheck(force: bool = False) -> str:
    """The launch-check variant: the product's engine AND its kernel translation units' host side
    (spmm.cu, dense.cu, gat.cu objects: launchers, kernel selection, grid arithmetic) on the fake runtime,
    which checks every launch against the hardware's launch limits and executes nothing.  The tcgen05
    launchers (driver-API tensor maps) and the communicator stay stubbed."""
    from dorylus_b200 import build as product_build

    product_build.build()
    out = os.path.join(OUT_DIR, "libdorylus_launchcheck.so")
    objs = [os.path.join(OBJ, n) for n in ("engine_cu.o", "spmm_cu.o", "dense_cu.o", "gat_cu.o", "loader_cpp.o", "partition_cpp.o", "tile_plan_cpp.o")]
    srcs = [os.path.join(HERE, n) for n in ("fake_cudart.cpp", "cpu_kernels.cpp")]
    deps = objs + srcs + [os.path.abspath(__file__)]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-DDORY_LAUNCHCHECK", "-Wl,-Bsymbolic", "-I" + CUDA_INC,
           "-o", out] + srcs + objs + ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("launch-check build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return out


def build_commcheck(force: bool = False) -> str:
    """The comm-check variant: the product's REAL communicator object (comm.cu: plans, staging, grouped
    all-to-all-v, the peer-memory store plan) over an NCCL whose ranks are threads of this process,
    pointer-carrying IPC handles and scalar statements of comm.cu's kernels (fake_nccl.cpp); compute as
    in the plain hostcheck build."""
    from dorylus_b200 import build as product_build

    product_build.build()
    out = os.path.join(OUT_DIR, "libdorylus_commcheck.so")
    objs = [os.path.join(OBJ, n) for n in ("engine_cu.o", "comm_cu.o", "loader_cpp.o", "partition_cpp.o", "tile_plan_cpp.o")]
    srcs = [os.path.join(HERE, n) for n in ("fake_cudart.cpp", "cpu_kernels.cpp", "fake_nccl.cpp")]
    deps = objs + srcs + [os.path.abspath(__file__)]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-DDORY_COMMCHECK", "-Wl,-Bsymbolic", "-I" + CUDA_INC,
           "-o", out] + srcs + objs + ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("comm-check build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return out


def build_driver(force: bool = False) -> str:
    """host/dorylus_b200_run.cpp linked against the hostcheck library instead of libdorylus_b200.so."""
    lib = build(force)
    src = os.path.join(ROOT, "host", "dorylus_b200_run.cpp")
    out = os.path.join(OUT_DIR, "dorylus_b200_run_hostcheck")
    deps = [src, lib, os.path.join(ROOT, "host", "saga_pipeline.hpp"), os.path.join(ROOT, "include", "dorylus_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", src, "-o", out, lib, "-Wl,-rpath," + OUT_DIR],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostcheck driver build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return out


if __name__ == "__main__":
    print(build(force=True))
