"""Host-side plan of the shared-memory-staged aggregation (dorylus_b200/csrc/tile_plan.cpp) on CPU: whatever the
parameters, the plan must be a regrouping of the SAME adjacency -- every row keeps its multiset of (source, weight)
pairs, the first group of a row lies inside its tile's window and the second outside, every row is owned by at
most one tile (exactly one unless its degree excludes it), tiles come heaviest first.  The kernels that consume
the plan are tested on the GPU (tests/test_gpu_tile.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "hostcheck", "_build")  # git-ignored build directory


class PlanSizes(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("tileRows", "windowRows", "nTiles", "nRows", "maxWrows", "maxTileRows")] + \
               [(n, C.c_uint64) for n in ("E", "inWindowEdges", "maxTileEdges")]


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libtileplan_test.so")
    srcs = [os.path.join(HERE, "native", "tile_plan_shim.cpp"), os.path.join(ROOT, "dorylus_b200", "csrc", "tile_plan.cpp")]
    deps = srcs + [os.path.join(ROOT, "dorylus_b200", "csrc", "tile_plan.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", so] + srcs + ["-lpthread"], check=True)
    lib = C.CDLL(so)
    lib.tp_build.restype = C.c_void_p
    lib.tp_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_uint32] * 8 + [C.c_int, C.c_double, C.POINTER(PlanSizes)]
    lib.tp_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.tp_free.argtypes = [C.c_void_p]
    lib.tp_estimate.restype = C.c_double
    lib.tp_estimate.argtypes = [C.c_void_p, C.c_void_p] + [C.c_uint32] * 5
    return lib


def community_graph(V, deg, communities, locality, seed, hubs=()):
    """CSC adjacency (offsets u64, sources u32, weights f32) with `locality` of every row's edges inside its
    community of consecutive ids; `hubs` = (row, degree) pairs appended on top."""
    rng = np.random.default_rng(seed)
    degs = np.maximum(1, rng.poisson(deg, V)).astype(np.int64)
    degs[rng.integers(0, V, V // 50)] = 0  # some empty rows
    for r, d in hubs:
        degs[r] = d
    ptrs = np.zeros(V + 1, np.uint64)
    ptrs[1:] = np.cumsum(degs)
    E = int(ptrs[-1])
    dst = np.repeat(np.arange(V), degs)
    csize = (V + communities - 1) // communities
    local = rng.random(E) < locality
    inside = (dst // csize) * csize + rng.integers(0, csize, E)
    idx = np.where(local, np.minimum(inside, V - 1), rng.integers(0, V, E)).astype(np.uint32)
    vals = rng.random(E, dtype=np.float32) + 0.25
    return ptrs, idx, vals


def build(lib, ptrs, idx, vals, V, nsrc, tileRows=0, windowRows=0, maxWindowRows=2048, teamDegree=512, excludeDegree=0,
          edgeCap=0, keepRowOrder=False, minTileCoverage=0.25):
    sz = PlanSizes()
    h = lib.tp_build(ptrs.ctypes.data, idx.ctypes.data, vals.ctypes.data, V, nsrc, tileRows, windowRows, maxWindowRows,
                     teamDegree, excludeDegree, edgeCap, int(keepRowOrder), minTileCoverage, C.byref(sz))
    try:
        shapes = {0: (np.uint64, 2 * V + 1), 1: (np.uint32, sz.E), 2: (np.float32, sz.E), 3: (np.uint32, sz.nRows),
                  4: (np.uint32, sz.nTiles + 1), 5: (np.uint32, sz.nTiles), 6: (np.uint32, sz.nTiles),
                  7: (np.uint32, sz.nTiles), 8: (np.uint64, sz.nTiles), 9: (np.uint64, sz.nTiles)}
        out = {}
        for which, (dt, n) in shapes.items():
            a = np.zeros(max(int(n), 1), dt)
            lib.tp_copy(h, which, a.ctypes.data)
            out[which] = a[:int(n)]
    finally:
        lib.tp_free(h)
    names = ["ptrs", "idx", "vals", "rows", "tilePtr", "tileTeam", "tileWlo", "tileWrows", "tileE0", "tileE1"]
    return sz, {names[k]: v for k, v in out.items()}


def check_plan(ptrs, idx, vals, V, nsrc, sz, p, excludeDegree=0, edgeCap=0, keepRowOrder=False, teamDegree=512):
    deg = (ptrs[1:] - ptrs[:-1]).astype(np.int64)
    E = int(ptrs[-1])
    assert sz.E == E and p["ptrs"][2 * V] == E
    pp = p["ptrs"].astype(np.int64)
    # every row keeps its edge range and its multiset of (source, weight); [b, mid) is the in-window group
    assert np.array_equal(pp[0:2 * V:2], ptrs[:-1].astype(np.int64))
    mid = pp[1:2 * V:2]
    assert np.all(mid >= pp[0:2 * V:2]) and np.all(mid <= ptrs[1:].astype(np.int64))
    key_old = np.lexsort((vals, idx, np.repeat(np.arange(V), deg)))
    key_new = np.lexsort((p["vals"], p["idx"], np.repeat(np.arange(V), deg)))
    assert np.array_equal(idx[key_old], p["idx"][key_new]) and np.array_equal(vals[key_old], p["vals"][key_new])
    # rows: each at most once; exactly the rows below the exclusion degree
    rows = p["rows"].astype(np.int64)
    assert len(np.unique(rows)) == len(rows)
    owned = np.zeros(V, bool)
    owned[rows] = True
    want = deg < excludeDegree if excludeDegree else np.ones(V, bool)
    assert np.array_equal(owned, want)
    assert np.all(mid[~owned] == pp[0:2 * V:2][~owned])  # rows outside every tile: everything in the "rest" group
    tp = p["tilePtr"].astype(np.int64)
    assert tp[0] == 0 and tp[-1] == len(rows) and np.all(np.diff(tp) > 0)
    assert int(np.diff(tp).max()) == sz.maxTileRows <= sz.tileRows
    tile_edges = (p["tileE1"] - p["tileE0"]).astype(np.int64)
    assert np.all(np.diff(tile_edges) <= 0), "tiles are issued heaviest first"
    assert int(tile_edges.max()) == sz.maxTileEdges
    in_window = 0
    for t in range(sz.nTiles):
        r = rows[tp[t]:tp[t + 1]]
        lo, n = int(p["tileWlo"][t]), int(p["tileWrows"][t])
        assert n <= sz.windowRows and lo + n <= nsrc
        assert r.max() - r.min() + 1 == len(r), "a tile is a run of consecutive rows"
        assert ptrs[r.min()] == p["tileE0"][t] and ptrs[r.max() + 1] == p["tileE1"][t]
        if edgeCap and len(r) > 1:
            assert tile_edges[t] <= edgeCap
        if keepRowOrder:
            assert np.array_equal(r, np.arange(r.min(), r.max() + 1))
        else:
            assert np.all(np.diff(deg[r]) <= 0), "degree-descending inside a tile"
            assert int(p["tileTeam"][t]) == int(np.sum(deg[r] >= teamDegree))
        for v in r:
            b, m, e = pp[2 * v], pp[2 * v + 1], pp[2 * v + 2]
            s_in, s_out = p["idx"][b:m].astype(np.int64), p["idx"][m:e].astype(np.int64)
            assert np.all((s_in >= lo) & (s_in < lo + n))
            assert np.all((s_out < lo) | (s_out >= lo + n))
            in_window += m - b
    assert in_window == sz.inWindowEdges
    return in_window / max(E, 1)


@pytest.mark.parametrize("case", [
    dict(),                                                    # automatic window and tile size
    dict(tileRows=32, windowRows=128),
    dict(tileRows=48, windowRows=96, keepRowOrder=True, edgeCap=600, excludeDegree=300),   # the low-degree kernel's plan
    dict(windowRows=64, minTileCoverage=0.9),                 # most tiles stage nothing
    dict(tileRows=64, windowRows=4096, maxWindowRows=4096),   # window larger than the source block
], ids=["auto", "fixed", "low-degree", "strict-coverage", "huge-window"])
def test_plan_is_a_regrouping_of_the_adjacency(lib, case):
    V = 3000
    ptrs, idx, vals = community_graph(V, deg=20, communities=40, locality=0.85, seed=5, hubs=((17, 900), (2001, 450)))
    kw = dict(teamDegree=400)
    kw.update(case)
    sz, p = build(lib, ptrs, idx, vals, V, V, **kw)
    cov = check_plan(ptrs, idx, vals, V, V, sz, p, excludeDegree=kw.get("excludeDegree", 0), edgeCap=kw.get("edgeCap", 0),
                     keepRowOrder=kw.get("keepRowOrder", False), teamDegree=kw["teamDegree"])
    if not case or case.get("windowRows", 0) >= 96 and case.get("minTileCoverage", 0) < 0.5:
        assert cov > 0.6, "75-row communities: a window of >= 96 rows holds most of a tile's local edges"


def test_ghost_rows_and_structureless_graph(lib):
    """Sources may lie beyond the local rows (ghost block: nSrcRows > V); without locality the best window holds
    about windowRows / nSrcRows of the edges and tiles below minTileCoverage stage nothing."""
    V, nsrc = 1500, 4000
    rng = np.random.default_rng(3)
    degs = rng.integers(0, 40, V)
    ptrs = np.zeros(V + 1, np.uint64)
    ptrs[1:] = np.cumsum(degs)
    E = int(ptrs[-1])
    idx = rng.integers(0, nsrc, E).astype(np.uint32)
    vals = rng.random(E, dtype=np.float32)
    sz, p = build(lib, ptrs, idx, vals, V, nsrc, tileRows=64, windowRows=256)
    cov = check_plan(ptrs, idx, vals, V, nsrc, sz, p)
    assert cov == 0.0 and not p["tileWrows"].any()
    sz, p = build(lib, ptrs, idx, vals, V, nsrc, tileRows=64, windowRows=256, minTileCoverage=0.0)
    cov = check_plan(ptrs, idx, vals, V, nsrc, sz, p)
    assert 0.04 < cov < 0.15
    est = lib.tp_estimate(ptrs.ctypes.data, idx.ctypes.data, V, nsrc, 64, 256, 1)
    assert abs(est - cov) < 1e-9, "the estimate over every tile is the plan's own coverage"


def test_empty_and_single_row(lib):
    ptrs = np.zeros(2, np.uint64)
    sz, p = build(lib, ptrs, np.zeros(1, np.uint32), np.zeros(1, np.float32), 1, 1)
    assert sz.E == 0 and sz.inWindowEdges == 0 and list(p["rows"]) == [0]
    ptrs = np.array([0, 3], np.uint64)
    idx = np.array([0, 0, 0], np.uint32)
    vals = np.array([1, 2, 3], np.float32)
    sz, p = build(lib, ptrs, idx, vals, 1, 1, minTileCoverage=0.0)
    assert sz.inWindowEdges == 3 and list(p["vals"]) == [1, 2, 3], "stable inside a group"
