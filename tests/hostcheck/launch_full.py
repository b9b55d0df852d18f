#!/usr/bin/env python
"""The launch check at the sizes BASELINE.json names (not collected by pytest; ~5 minutes, ~20 GB of
host memory, CPU only):  python tests/hostcheck/launch_full.py

Builds the Reddit shape, one eighth of the Friendster shape and the Amazon shape (whole, and partition
3 of 8 with its ghost blocks), loads each into the engine on the launch-check runtime
(tests/hostcheck/build.py: build_launchcheck -- the product's engine and kernel launchers, every launch
checked against the sm_100 limits, nothing executed, device memory never touched) and runs one epoch
per schedule (GCN reference order, GCN apply-first, GAT with source windows).  This is where the
V-dependent size of the GAT edge-backward workspace was found to be missing.
"""
import ctypes as C
import importlib.util
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from dorylus_b200 import _lib  # noqa: E402

spec_ = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(HERE, "build.py"))
mod = importlib.util.module_from_spec(spec_)
spec_.loader.exec_module(mod)
lib = C.CDLL(mod.build_launchcheck())
for name, (res, args) in _lib.SYMBOLS.items():
    fn = getattr(lib, name)
    fn.restype, fn.argtypes = res, args
_lib._lib = lib

from dorylus_b200 import engine as dengine  # noqa: E402
from dorylus_b200 import synth  # noqa: E402
from dorylus_b200.engine import BACKWARD, FORWARD, GAT, GCN, Engine  # noqa: E402


def log():
    n, v = C.c_ulonglong(), C.c_ulonglong()
    buf = C.create_string_buffer(512)
    lib.hostcheck_launch_log(C.byref(n), C.byref(v), buf, 512)
    return int(n.value), int(v.value), buf.value.decode()


def whole(name, spec, gat=True):
    t0 = time.time()
    src, dst = synth.generate_edges(spec)
    image = dengine.preprocess_edges(src, dst, np.zeros(spec.num_vertices, np.int32), spec.num_vertices, 0, 1)
    print("%s: V=%d E=%d built in %.0fs" % (name, spec.num_vertices, src.size, time.time() - t0), flush=True)
    for label, flags, gnn in (("GCN reference order", 0, GCN), ("GCN apply-first", _lib.FLAG_APPLY_FIRST, GCN),
                              ("GAT + source windows", _lib.FLAG_GAT_PREDICT_AH, GAT)):
        if gnn == GAT and not gat:
            continue
        e = Engine(spec.dims, gnn, flags=flags)
        if gnn == GAT:
            e.set_option("gat_windows", 1)
        e.load_partition(image)
        with e:
            e.init_weights()
            e.epoch()
        n, bad, first = log()
        print("  %-22s launches %3d  violations %d %s" % (label, n, bad, first), flush=True)
        assert bad == 0
    return src, dst


def main():
    whole("reddit", synth.CONFIGS["reddit"])
    whole("friendster/8", synth.GraphSpec("friendster/8", 8201045, 225000000, [16, 48, 51], seed=77, sigma=0.9, locality=0.9,
                                          communities=4096))
    spec = synth.CONFIGS["amazon"]
    src, dst = whole("amazon", spec, gat=False)
    parts = synth.contiguous_parts(spec.num_vertices, 8)
    img = dengine.preprocess_edges(src, dst, parts, spec.num_vertices, 3, 8)
    del src, dst
    L = len(spec.dims) - 1
    for label, flags in (("reference order", 0), ("apply-first", _lib.FLAG_APPLY_FIRST)):
        e = Engine(spec.dims, GCN, node_id=3, num_nodes=8, flags=flags)
        e.load_partition(img)
        with e:
            e.init_weights()
            sched = [e.apply_first(l) for l in range(L)]
            for l in range(L):  # the operators of one epoch without the exchanges (no communicator here)
                c = e.whole_chunk(l, FORWARD)
                if sched[l]:
                    e.applyVertex(c)
                    e.aggregate(c)
                else:
                    e.aggregate(c)
                    e.applyVertex(c)
            for l in list(range(L - 1, 0, -1)) + ([0] if sched[0] else []):
                c = e.whole_chunk(l, BACKWARD)
                e.aggregate(c)
                e.applyVertex(c)
            n, bad, first = log()
            print("amazon partition 3 of 8 (%d + %d ghost rows), %s: launches %d violations %d %s"
                  % (e.srcGhostCnt, e.dstGhostCnt, label, n, bad, first), flush=True)
            assert bad == 0


if __name__ == "__main__":
    main()
