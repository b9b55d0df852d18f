#!/usr/bin/env python
"""Fuzzers for the engine's host logic on the emulated runtime (not collected by pytest; minutes, CPU only).

    python tests/hostcheck/fuzz.py [options|images|host|all]

    # with AddressSanitizer (the scalar kernel statements, loader.cpp and partition.cpp instrumented):
    g++ -std=c++17 -O1 -g -fPIC -shared -w -fsanitize=address -Wl,-Bsymbolic -I/usr/local/cuda/include \\
        -o /tmp/libhostcheck_asan.so tests/hostcheck/fake_cudart.cpp tests/hostcheck/cpu_kernels.cpp \\
        dorylus_b200/csrc/loader.cpp dorylus_b200/csrc/partition.cpp dorylus_b200/_obj/engine_cu.o -lpthread
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:allocator_may_return_null=1 \\
        DORY_HOSTCHECK_LIB=/tmp/libhostcheck_asan.so python tests/hostcheck/fuzz.py all

  options  random tuning options (windows, heavy / hub thresholds, row orders, kernel shapes), partitions
           with ghosts, whole and sub-range chunks: the aggregation against the oracle
  images   truncated / bit-flipped / count-corrupted graph.<id>.bin images into dory_load_partition and
           dory_ghost_slots: an error code or a valid load, never a crash
  host     dory_preprocess_edges / dory_partition_edges on out-of-range ids and owners, dory_read_features /
           dory_read_labels on damaged files
(The call-sequence fuzzer over the operators is a regular test: test_abi_survives_arbitrary_call_sequences.)
"""
import ctypes as C
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from dorylus_b200 import _lib  # noqa: E402


def load_hostcheck():
    path = os.environ.get("DORY_HOSTCHECK_LIB")
    if not path:
        spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        path = mod.build()
    lib = C.CDLL(path)
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib._lib = lib


def fuzz_options():
    from helpers import random_dataset, rel_err
    from dorylus_b200.engine import GCN, GAT, Engine, Chunk, DoryError, FORWARD, BACKWARD
    from oracle.pyoracle import Oracle
    o = Oracle()
    rng = np.random.default_rng(3)
    bad = 0
    for trial in range(80):
        P = int(rng.integers(1, 4))
        F = int(rng.choice([3, 16, 17, 48, 100, 130, 200]))
        dims = [F, int(rng.integers(2, 20)), int(rng.integers(2, 9))]
        V = int(rng.integers(P * 3, 400)); E = int(rng.integers(0, 12 * V))
        hub = (np.concatenate([np.arange(1, min(V, 200)), np.zeros(min(V, 200) - 1, np.int64)]), np.concatenate([np.zeros(min(V, 200) - 1, np.int64), np.arange(1, min(V, 200))])) if rng.random() < 0.5 else None
        ds = random_dataset(V=V, E_und=max(E, 1), dims=dims, P=P, seed=trial, sigma=float(rng.choice([0.5, 1.0, 1.5])), extra_edges=hub)
        p = int(rng.integers(0, P)); g = ds.graphs[p]
        if g.local_vtx_cnt == 0: continue
        e = Engine(dims, GCN, node_id=p, num_nodes=P)
        opts = {}
        for k, vals in (("src_blocks", [0, 1, 2, 5, 64, 100]), ("heavy_degree", [0, 1, 8, 64, 1024]), ("hub_degree", [0, 1, 16, 100, 5000]),
                        ("row_order", [0, 1, 2]), ("locality_block", [0, 1, 7, 1000]), ("spmm_light", [0, 1, 2]), ("spmm_lg", [0, 4, 8, 16, 32]),
                        ("spmm_vec", [0, 1, 2, 4]), ("spmm_unroll", [0, 1, 2]), ("spmm_occ", [0, 4, 6, 8])):
            if rng.random() < 0.5:
                v = int(rng.choice(vals)); opts[k] = v
                try: e.set_option(k, v)
                except DoryError: pass
        try:
            e.load_partition(ds.images[p])
        except DoryError as ex:
            print("load error", trial, ex); e.close(); continue
        with e:
            x = ds.feats[g.local_to_global]; xg = ds.feats[g.src_ghost_gvid]
            e.set_tensor(0, "x", x)
            if g.src_ghost_cnt: e.set_tensor(0, "fg", xg)
            lo = int(rng.integers(0, g.local_vtx_cnt)); up = int(rng.integers(lo, g.local_vtx_cnt + 1))
            for (low, upp) in ((0, g.local_vtx_cnt), (lo, up)):
                e.set_tensor(0, "ah", np.zeros((g.local_vtx_cnt, F), np.float32))
                e.aggregate(Chunk(0, p, low, upp, 0, FORWARD, 1, True))
                want = o.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, x, xg if g.src_ghost_cnt else None)
                got = e.get_tensor(0, "ah")
                err = rel_err(got[low:upp], want[low:upp]) if upp > low else 0.0
                clean = not got[:low].any() and not got[upp:].any()
                if err > 1e-5 or not clean:
                    bad += 1; print("BAD", trial, dims, "V", V, "P", P, opts, (low, upp), err, clean)
    print("done, bad", bad)


def fuzz_images():
    from helpers import random_dataset
    from dorylus_b200 import engine as dengine
    from dorylus_b200.engine import GCN, Engine, DoryError
    rng = np.random.default_rng(0)
    ds = random_dataset(V=60, E_und=300, dims=[8, 4, 3], P=2, seed=1)
    good = np.frombuffer(ds.images[0], dtype=np.uint8)
    other = ds.images[1]
    n_err = n_ok = 0
    for trial in range(3000):
        img = good.copy()
        mode = trial % 4
        if mode == 0:   # truncate
            img = img[: int(rng.integers(0, img.size))]
        elif mode == 1:  # flip a few bytes anywhere
            for _ in range(int(rng.integers(1, 4))):
                img[int(rng.integers(0, img.size))] = int(rng.integers(0, 256))
        elif mode == 2:  # corrupt the header counts
            img[int(rng.integers(0, 44))] = int(rng.integers(0, 256))
        else:            # huge counts
            pos = int(rng.integers(0, 10)) * 4
            img[pos:pos + 4] = np.frombuffer(np.uint32(rng.integers(0, 2**32)).tobytes(), dtype=np.uint8)
        b = img.tobytes()
        try:
            dengine.ghost_slots(b, 0, other, 0); dengine.ghost_slots(other, 1, b, 1); n_ok += 1
        except DoryError:
            n_err += 1
        e = Engine([8, 4, 3], GCN, node_id=0, num_nodes=2)
        try:
            e.load_partition(b); n_ok += 1
        except DoryError:
            n_err += 1
        finally:
            e.close()
    print("no crash: ok", n_ok, "errors", n_err)


def fuzz_host():
    from dorylus_b200 import engine as dengine, formats
    from dorylus_b200.engine import DoryError
    rng = np.random.default_rng(0)
    ok = err = 0
    for trial in range(400):
        V = int(rng.integers(0, 50)); E = int(rng.integers(0, 200)); P = int(rng.integers(1, 5))
        hi = max(V, 1) + (int(rng.integers(0, 5)) if rng.random() < 0.3 else 0)   # sometimes ids beyond V
        src = rng.integers(0, hi, E).astype(np.uint32); dst = rng.integers(0, hi, E).astype(np.uint32)
        parts = rng.integers(-1 if rng.random() < 0.2 else 0, P + (2 if rng.random() < 0.2 else 0), max(V, 0)).astype(np.int32)
        for p in range(P):
            try:
                img = dengine.preprocess_edges(src, dst, parts, V, p, P, bool(trial & 1)); ok += 1
                formats.parse_graph_bin(img)
            except DoryError: err += 1
        # partitioner
        try:
            dengine.partition_edges(src, dst, V, int(rng.integers(0, 6)), passes=int(rng.integers(0, 3))); ok += 1
        except DoryError: err += 1
    # dataset readers on damaged files
    d = tempfile.mkdtemp() + "/"
    V, F, K = 30, 7, 4
    src = np.arange(V, dtype=np.uint32); img = dengine.preprocess_edges(src, (src + 1) % V, np.zeros(V, np.int32), V, 0, 1)
    feats = rng.random((V, F), dtype=np.float32); labels = rng.integers(0, K, V).astype(np.uint32)
    for trial in range(300):
        formats.write_features(d + "f.bsnap", feats); formats.write_labels(d + "l.bsnap", labels, K)
        for name in ("f.bsnap", "l.bsnap"):
            b = bytearray(open(d + name, "rb").read())
            m = trial % 3
            if m == 0: b = b[: int(rng.integers(0, len(b)))]
            elif m == 1: b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            else: b[0:4] = np.uint32(rng.integers(0, 2**32)).tobytes()
            open(d + name, "wb").write(bytes(b))
        for f in os.listdir(d):
            if f.startswith("feats"): os.remove(d + f)
        try: dengine.read_features(d, d + "f.bsnap", img, 0, F); ok += 1
        except DoryError: err += 1
        try: dengine.read_labels(d + "l.bsnap", img, K); ok += 1
        except DoryError: err += 1
    print("no crash: ok", ok, "errors", err)


if __name__ == "__main__":
    load_hostcheck()
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("options", "all"):
        fuzz_options()
    if which in ("images", "all"):
        fuzz_images()
    if which in ("host", "all"):
        fuzz_host()
