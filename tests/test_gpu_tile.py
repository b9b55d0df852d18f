"""GPU parity of the shared-memory-staged aggregation (csrc/spmm_tile.cu + tile_plan.cpp) against the
CPU oracle: graphs with community structure in their vertex numbering, high-degree (warp / CTA per
row, column slabs) and low-degree (lane group per row) modes, forced windows smaller and larger than a
community, rows excluded from the tiles, whole epochs in both schedules.  Bar: 1e-5."""
import numpy as np
import pytest

from dorylus_b200 import _lib, formats, synth
from dorylus_b200 import engine as dengine
from dorylus_b200.engine import BACKWARD, FORWARD, GCN, Engine
from helpers import rel_err
from oracle.driver import OracleGCN

pytestmark = pytest.mark.gpu
TOL = 1e-5


class Community:
    def __init__(self, V, deg, dims, communities, locality, seed, sigma=0.9, extra=None):
        spec = synth.GraphSpec("c", V, V * deg, list(dims), seed=seed, sigma=sigma, locality=locality, communities=communities)
        src, dst, _, _ = synth.generate_incident_edges(spec, 0, 1)
        if extra is not None:
            src = np.concatenate([src, extra[0].astype(np.uint32)])
            dst = np.concatenate([dst, extra[1].astype(np.uint32)])
        self.V, self.dims = V, list(dims)
        self.image = dengine.preprocess_edges(src, dst, np.zeros(V, np.int32), V, 0, 1)
        self.graph = formats.parse_graph_bin(self.image)
        self.feats = synth.generate_features(V, dims[0], seed + 1)
        self.onehot = formats.one_hot(synth.generate_labels(V, dims[-1], seed + 2), dims[-1])


def engine_for(ds, opts, flags=0):
    e = Engine(ds.dims, GCN, flags=flags)
    for k, v in opts.items():
        e.set_option(k, v)
    e.load_partition(ds.image)
    e.set_tensor(0, "x", ds.feats)
    e.set_tensor(len(ds.dims) - 2, "lab", ds.onehot)
    e.init_weights()
    return e


def check_aggregations(oracle, ds, e):
    g = ds.graph
    e.aggregate(e.whole_chunk(0, FORWARD))
    want = oracle.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, ds.feats, None)
    assert rel_err(e.get_tensor(0, "ah"), want) < TOL
    grad = np.random.default_rng(3).standard_normal((ds.V, ds.dims[1])).astype(np.float32)
    e.set_tensor(1, "grad", grad)
    e.aggregate(e.whole_chunk(1, BACKWARD))
    want_b = oracle.aggregate_gcn(g.row_ptrs, g.col_idxs, g.bwd_vals, g.norms, grad, None)
    assert rel_err(e.get_tensor(0, "aTg"), want_b) < TOL


HUB = (np.concatenate([np.arange(1, 1500), np.zeros(1499, np.int64)]),
       np.concatenate([np.zeros(1499, np.int64), np.arange(1, 1500)]))  # vertex 0: degree >= 1499

LOW = {
    "friendster-widths": dict(V=20000, deg=24, dims=[16, 48, 51], communities=80, locality=0.9, seed=7),
    "amazon-widths": dict(V=12000, deg=24, dims=[100, 64, 64, 25], communities=40, locality=0.9, seed=9),
    "narrow+hub": dict(V=9000, deg=20, dims=[8, 12, 5], communities=30, locality=0.85, seed=11, extra=HUB),
}


@pytest.mark.parametrize("opts", [{}, {"tile_window": 128, "tile_rows": 50}, {"tile_window": 1024, "tile_rows": 333},
                                  {"tile_smem_kb": 8}, {"tile_pipe": 0}, {"tile_pipe": 0, "tile_window": 128, "tile_edges": 1024}],
                         ids=["auto", "small-window", "large-window", "8KB", "one-tile-per-cta", "one-tile-per-cta-small"])
@pytest.mark.parametrize("name", list(LOW))
def test_low_degree_tiles(oracle, name, opts):
    ds = Community(**LOW[name])
    with engine_for(ds, dict(tile=1, **opts)) as e:
        info = e.tile_info(FORWARD)
        # 8 KB of shared memory holds 32 rows of the widest layer: no tile finds a window worth staging and
        # every edge takes the kernel's L2 path (coverage 0) -- still the tile kernel, still the same numbers
        assert info["n_tiles"] > 0 and (info["coverage"] > 0.2 or "tile_smem_kb" in opts), info
        check_aggregations(oracle, ds, e)
        orc = OracleGCN(oracle, [ds.graph], ds.dims)
        orc.load_features(ds.feats, ds.onehot)
        for ep in range(2):
            want = orc.epoch()
            st = e.epoch()
            L = len(ds.dims) - 1
            for l in range(L):
                assert rel_err(e.get_tensor(l, "ah"), orc.saved[0][l]["ah"]) < TOL, (ep, l)
                assert rel_err(e.get_weight_grad(l), orc.dW[0][l]) < TOL, (ep, l)
                e.set_weights(l, orc.W[l])
            assert st["acc_sum"] == want["acc"][0]


HIGH = dict(V=4096, deg=160, dims=[602, 128, 41], communities=32, locality=0.8, seed=13, sigma=1.0)


@pytest.mark.parametrize("opts", [{"tile_slab": 64}, {"tile_slab": 128, "tile_team": 64}, {"tile_slab": 32, "tile_window": 96, "tile_rows": 40},
                                  {"tile_slab": 96, "tile_window": 256, "tile_rows": 128}],
                         ids=["slab64", "slab128-team64", "slab32-small", "slab96"])
@pytest.mark.parametrize("apply_first", [False, True], ids=["reference-order", "apply-first"])
def test_high_degree_tiles(oracle, opts, apply_first):
    ds = Community(**HIGH, extra=HUB)
    with engine_for(ds, dict(tile=1, **opts), flags=_lib.FLAG_APPLY_FIRST if apply_first else 0) as e:
        info = e.tile_info(BACKWARD)
        assert info["n_tiles"] > 0 and info["coverage"] > 0.3, info
        if not apply_first:
            check_aggregations(oracle, ds, e)
        orc = OracleGCN(oracle, [ds.graph], ds.dims)
        orc.load_features(ds.feats, ds.onehot)
        want = orc.epoch()
        st = e.epoch()
        t = orc.saved[0]
        assert rel_err(e.get_tensor(0, "h"), t[0]["h"]) < TOL
        assert rel_err(e.get_tensor(0, "aTg"), t[0]["aTg"]) < (2 * TOL if apply_first else TOL)
        for l in range(2):
            assert rel_err(e.get_weight_grad(l), orc.dW[0][l]) < (2 * TOL if apply_first else TOL), l
        assert st["acc_sum"] == want["acc"][0]


def test_tile_results_are_bit_reproducible_and_auto_mode_declines_without_locality(oracle):
    ds = Community(**LOW["friendster-widths"])
    with engine_for(ds, dict(tile=1)) as e:
        c = e.whole_chunk(0, FORWARD)
        e.aggregate(c)
        a = e.get_tensor(0, "ah")
        e.aggregate(c)
        assert np.array_equal(a, e.get_tensor(0, "ah"))
    spec = synth.CONFIGS["reddit-small"]  # Chung-Lu, no communities: the default mode keeps the gather kernels
    src, dst = synth.generate_edges(spec)
    image = dengine.preprocess_edges(src, dst, np.zeros(spec.num_vertices, np.int32), spec.num_vertices, 0, 1)
    with Engine(spec.dims, GCN) as e:  # default mode (2): no locality, no plan
        e.load_partition(image)
        assert e.tile_info(FORWARD)["n_tiles"] == 0
    with Engine(ds.dims, GCN) as e:  # default mode on a LOW-degree graph with locality: needs an explicit tile=1
        e.load_partition(ds.image)
        assert e.tile_info(FORWARD)["n_tiles"] == 0


def test_default_mode_stages_a_high_degree_graph_with_locality(oracle):
    """Default options on the community-structured high-degree graph: the plan is kept (coverage >= 50 %), the
    per-launch slab choice runs (96-float slabs at F = 602, 64-float slabs at F = 128) and the epoch matches."""
    ds = Community(**HIGH)
    with engine_for(ds, {}) as e:
        info = e.tile_info(FORWARD)
        assert info["n_tiles"] > 0 and info["coverage"] >= 0.5, info
        check_aggregations(oracle, ds, e)
        orc = OracleGCN(oracle, [ds.graph], ds.dims)
        orc.load_features(ds.feats, ds.onehot)
        want = orc.epoch()
        st = e.epoch()
        assert rel_err(e.get_tensor(0, "h"), orc.saved[0][0]["h"]) < TOL
        for l in range(2):
            assert rel_err(e.get_weight_grad(l), orc.dW[0][l]) < TOL, l
        assert st["acc_sum"] == want["acc"][0]
