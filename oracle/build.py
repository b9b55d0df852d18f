"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Two artefacts, both plain g++ invocations (the reference's own CMake build is NOT run):

* ``oracle/_build/liboracle.so``  -- our line-faithful restatement (oracle/oracle.cpp), built with
  the flags of the reference's CPU backend: ``-O3 -march=native -fopenmp``
  (reference CMakeLists.txt:7, ``_CPU_ENABLED_`` pragma at engine/ops/gcn_ops.cpp:159-161).
* ``oracle/_ref/libdoryref.so``   -- the reference's own loader / Matrix / Adam translation units
  compiled *where they lie* under /root/reference (never copied), plus oracle/ref_driver.cpp.
  Only buildable where /root/reference exists; the prebuilt .so travels to the GPU box.

Nothing in the product path imports this module.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DORYLUS_REFERENCE", "/root/reference")
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")

REF_SOURCES = [
    "src/common/matrix.cpp",
    "src/common/utils.cpp",
    "src/graph-server/graph/graph.cpp",
    "src/graph-server/graph/vertex.cpp",
    "src/graph-server/graph/edge.cpp",
    "src/graph-server/graph/dataloader.cpp",
    "src/graph-server/utils/utils.cpp",
    "src/weight-server/AdamOptimizer.cpp",
    "src/weight-server/weighttensor.cpp",
]
# #included (inside namespaces) by ref_driver.cpp: the Lambda functions' tensor ops
REF_INCLUDED = [
    "src/funcs/gcn/ops/forward_ops.cpp",
    "src/funcs/gcn/ops/backward_ops.cpp",
    "src/funcs/gat/ops/forward_ops.cpp",
    "src/funcs/gat/ops/backward_ops.cpp",
]


def openblas_path() -> str:
    """The LP64 OpenBLAS bundled with scipy (exports scipy_cblas_sgemm)."""
    import scipy  # noqa: F401  (only to locate site-packages)

    libs = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(libs, "libscipy_openblas-*.so")))
    if not hits:
        raise RuntimeError("scipy's bundled OpenBLAS not found under %s" % libs)
    return hits[0]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("oracle build failed")


def build_oracle(force: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    out = os.path.join(BUILD_DIR, "liboracle.so")
    src = os.path.join(HERE, "oracle.cpp")
    if force or _newer(out, [src, os.path.join(HERE, "shim", "cblas.h"), __file__]):
        blas = openblas_path()
        _run(["g++", "-std=c++14", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared",
              "-I" + os.path.join(HERE, "shim"), src, "-o", out,
              blas, "-Wl,-rpath," + os.path.dirname(blas)])
    return out


def build_ref(force: bool = False) -> str | None:
    """Compile the reference translation units in place.  Returns None when /root/reference is
    absent and no prebuilt library exists (e.g. a fresh GPU box without the snapshot)."""
    out = os.path.join(REF_DIR, "libdoryref.so")
    if not os.path.isdir(REF_ROOT):
        return out if os.path.exists(out) else None
    os.makedirs(REF_DIR, exist_ok=True)
    srcs = [os.path.join(REF_ROOT, s) for s in REF_SOURCES]
    drv = os.path.join(HERE, "ref_driver.cpp")
    if force or _newer(out, srcs + [drv, __file__] + [os.path.join(REF_ROOT, s) for s in REF_INCLUDED]):
        blas = openblas_path()
        objs = []
        for i, s in enumerate(srcs + [drv]):
            o = os.path.join(REF_DIR, "obj%d_%s.o" % (i, os.path.basename(s).replace(".cpp", "")))
            # -O2 as in the survey probe; -march=native -O3 is the reference's Release flag set but
            # the loader is integer work and Matrix::dot is a BLAS call, so the level is immaterial.
            _run(["g++", "-std=c++11", "-O2", "-fPIC", "-w", "-I" + os.path.join(HERE, "shim"),
                  "-I" + os.path.join(REF_ROOT, "src"), "-c", s, "-o", o])
            objs.append(o)
        _run(["g++", "-shared", "-o", out] + objs +
             [blas, "-Wl,-rpath," + os.path.dirname(blas), "-lpthread"])
    return out


# The reference's hot path itself: Engine::aggregateGCN / preallocateGCN (engine/ops/gcn_ops.cpp) and CPUComm
# (commmanager/CPU_comm.cpp: vtxNNForwardGCN / vtxNNBackwardGCN and their helpers), compiled in place against
# the stub <zmq.hpp> / Boost headers of oracle/shim/ and linked with oracle/ref_engine.cpp (an in-process
# stand-in for the ZeroMQ weight-server client; see its header).  Flags of the reference's CPU backend:
# -O3 -march=native -fopenmp -D_CPU_ENABLED_ (CMakeLists.txt:7,23).
REF_ENGINE_SOURCES = [
    "src/graph-server/engine/ops/gcn_ops.cpp",
    "src/graph-server/engine/ops/gat_ops.cpp",
    "src/graph-server/engine/ops/tensors.cpp",
    "src/graph-server/commmanager/CPU_comm.cpp",
    "src/common/matrix.cpp",
    "src/common/utils.cpp",
    "src/graph-server/graph/graph.cpp",
    "src/graph-server/graph/vertex.cpp",
    "src/graph-server/graph/edge.cpp",
    "src/graph-server/utils/utils.cpp",
]


def build_ref_engine(force: bool = False) -> str | None:
    out = os.path.join(REF_DIR, "librefengine.so")
    if not os.path.isdir(REF_ROOT):
        return out if os.path.exists(out) else None
    os.makedirs(REF_DIR, exist_ok=True)
    srcs = [os.path.join(REF_ROOT, s) for s in REF_ENGINE_SOURCES]
    drv = os.path.join(HERE, "ref_engine.cpp")
    shim = os.path.join(HERE, "shim")
    deps = srcs + [drv, __file__] + glob.glob(os.path.join(shim, "**", "*.h*"), recursive=True)
    if force or _newer(out, deps):
        blas = openblas_path()
        objs = []
        for i, s in enumerate(srcs + [drv]):
            o = os.path.join(REF_DIR, "eng%d_%s.o" % (i, os.path.basename(s).replace(".cpp", "")))
            _run(["g++", "-std=c++11", "-O3", "-march=native", "-fopenmp", "-fPIC", "-w", "-D_CPU_ENABLED_",
                  "-I" + shim, "-I" + os.path.join(REF_ROOT, "src"), "-c", s, "-o", o])
            objs.append(o)
        _run(["g++", "-shared", "-fopenmp", "-o", out] + objs + [blas, "-Wl,-rpath," + os.path.dirname(blas), "-lpthread",
                                                                   "-Wl,--no-undefined"])
    return out


# The reference's dataset tools that build from one file each (inputs/Makefile; labelsToBinary.cpp and
# featuresToBinary.cpp need Boost and do not).
REF_TOOLS = {
    "graphToBinary": "inputs/graphToBinary.cpp",      # text edge list -> graph.bsnap
    "generateFeatues": "inputs/generateFeatues.cpp",  # random features.bsnap writer ("<name>.feats")
    "generateLabels": "inputs/generateLabels.cpp",    # random labels.bsnap writer ("<name>.labels")
}


def build_ref_tools(force: bool = False) -> dict:
    """Compile the reference's own input tools in place into oracle/_ref/<tool>.  Returns
    {name: path} for the tools that exist (prebuilt ones included when /root/reference is absent)."""
    out = {}
    for name, rel in REF_TOOLS.items():
        exe = os.path.join(REF_DIR, name)
        src = os.path.join(REF_ROOT, rel)
        if os.path.exists(src):
            os.makedirs(REF_DIR, exist_ok=True)
            if force or _newer(exe, [src, __file__]):
                _run(["g++", "-std=c++11", "-O2", "-w", "-pthread", src, "-o", exe])
        if os.path.exists(exe):
            out[name] = exe
    return out


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
    print(build_ref_tools(force="--force" in sys.argv))
    print(build_ref_engine(force="--force" in sys.argv))
