"""Build recipe for libdorylus_b200.so (in-tree, sm_100a only).

    python -m dorylus_b200.build [--force] [--verbose]

Every translation unit is compiled with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (cross-compiles without a GPU) and the
objects are linked into ``dorylus_b200/libdorylus_b200.so``; the .so is git-ignored but travels to
the GPU box with the snapshot.  No other architecture is emitted: there is no multi-backend
dispatch and no CPU fallback.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libdorylus_b200.so")

SOURCES = ["engine.cu", "spmm.cu", "spmm_tile.cu", "dense.cu", "gat.cu", "gemm_tc.cu", "comm.cu", "loader.cpp", "partition.cpp",
           "tile_plan.cpp"]
HEADERS = ["common.cuh", "gat.cuh", "gemm_tc.cuh", "comm.h", "loader.h", "partition.h", "tile_plan.h", "../../include/dorylus_b200.h"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str, verbose: bool) -> tuple[str, str]:
    obj = os.path.join(OBJ, src.replace(".", "_") + ".o")
    cmd = [NVCC] + ARCH + CFLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:  # ptxas -v: registers / spills / smem per kernel
        f.write(log)
    if verbose:
        sys.stderr.write(log)
    return obj, log


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    todo, objs = [], []
    for s in SOURCES:
        obj = os.path.join(OBJ, s.replace(".", "_") + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, s)] + hdrs):
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    build_host_driver()
    return LIB


def build_host_driver() -> str:
    """host/dorylus_b200_run: the C++ driver over the C ABI (plain g++, links the .so above)."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "host", "dorylus_b200_run.cpp")
    out = os.path.join(root, "host", "dorylus_b200_run")
    if os.path.exists(src) and _stale(out, [src, LIB, os.path.join(root, "include", "dorylus_b200.h"),
                                            os.path.join(root, "host", "saga_pipeline.hpp")]):
        cmd = ["g++", "-std=c++17", "-O2", "-Wall", src, "-o", out, LIB, "-Wl,-rpath,$ORIGIN/../dorylus_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host driver build failed:\n%s\n%s" % (r.stdout, r.stderr))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
