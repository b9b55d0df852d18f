"""Parity at BASELINE.json's full size (Reddit shape: 232,965 vertices, 114.6 M edges) through
size-independent properties -- the CPU oracle needs ~3 s per aggregation here on 16 cores, so the
exact comparison is done on a row sample and the whole output is checked through identities:

  * row sums:   A_hat . 1 = norm[v] + sum of the edge weights of v          (forward and backward)
  * adjointness: <A_hat h, g> = <h, A_hat^T g>   ties the CSC forward to the CSR backward walk
  * linearity under an exact power-of-two scaling, bit reproducibility
  * oracle on a sample of destination rows (chunk sub-range of the reference API)
"""
import numpy as np
import pytest

from helpers import rel_err
from dorylus_b200 import engine as dengine
from dorylus_b200 import formats, synth
from dorylus_b200.engine import BACKWARD, FORWARD, GCN, Chunk, Engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reddit():
    spec = synth.CONFIGS["reddit"]
    src, dst = synth.generate_edges(spec)
    image = dengine.preprocess_edges(src, dst, np.zeros(spec.num_vertices, np.int32), spec.num_vertices, 0, 1)
    g = formats.parse_graph_bin(image)
    e = Engine(spec.dims, GCN)
    e.load_partition(image)
    yield spec, g, e
    e.close()


def test_counts_and_indexing_invariants(reddit):
    spec, g, e = reddit
    assert e.localVtxCnt == spec.num_vertices and e.localInEdgeCnt == spec.num_edges == e.localOutEdgeCnt
    assert g.col_ptrs[-1] == spec.num_edges and g.row_ptrs[-1] == spec.num_edges
    deg = np.diff(g.col_ptrs).astype(np.float64)
    assert np.allclose(g.norms, 1.0 / (deg + 1.0), rtol=1e-6)
    # the graph is symmetric (both directions present): in-degree == out-degree per vertex
    assert np.array_equal(np.diff(g.col_ptrs), np.diff(g.row_ptrs))
    assert int(deg.max()) >= 1024  # exercises the CTA-per-row kernel


def test_row_sums_adjointness_linearity_at_full_size(reddit, oracle):
    spec, g, e = reddit
    V, F = spec.num_vertices, spec.dims[1]
    rng = np.random.default_rng(9)
    # --- row sums, forward (layer 1 walks h[0]) and backward (grad[1])
    ones = np.ones((V, F), np.float32)
    e.set_tensor(0, "h", ones)
    e.aggregateGCN(e.whole_chunk(1, FORWARD))
    dst_of_edge = np.repeat(np.arange(V), np.diff(g.col_ptrs).astype(np.int64))
    rowsum_f = g.norms.astype(np.float64) + np.bincount(dst_of_edge, weights=g.fwd_vals.astype(np.float64), minlength=V)
    ah = e.get_tensor(1, "ah")
    assert rel_err(ah[:, 0], rowsum_f) < 1e-5 and rel_err(ah[:, F - 1], rowsum_f) < 1e-5
    e.set_tensor(1, "grad", ones)
    e.aggregateGCN(e.whole_chunk(1, BACKWARD))
    src_of_edge = np.repeat(np.arange(V), np.diff(g.row_ptrs).astype(np.int64))
    rowsum_b = g.norms.astype(np.float64) + np.bincount(src_of_edge, weights=g.bwd_vals.astype(np.float64), minlength=V)
    assert rel_err(e.get_tensor(0, "aTg")[:, 3], rowsum_b) < 1e-5
    # --- adjointness <A h, y> == <h, A^T y>
    h = rng.standard_normal((V, F)).astype(np.float32)
    y = rng.standard_normal((V, F)).astype(np.float32)
    e.set_tensor(0, "h", h)
    e.set_tensor(1, "grad", y)
    c_f, c_b = e.whole_chunk(1, FORWARD), e.whole_chunk(1, BACKWARD)
    e.aggregateGCN(c_f)
    e.aggregateGCN(c_b)
    ah, aTg = e.get_tensor(1, "ah"), e.get_tensor(0, "aTg")
    lhs = float(np.sum(ah.astype(np.float64) * y))
    rhs = float(np.sum(h.astype(np.float64) * aTg))
    scale = float(np.sqrt(np.sum(ah.astype(np.float64) ** 2) * np.sum(y.astype(np.float64) ** 2)))
    assert abs(lhs - rhs) <= 1e-6 * scale
    # --- bit reproducibility and exact scaling
    e.aggregateGCN(c_f)
    assert np.array_equal(e.get_tensor(1, "ah"), ah)
    e.set_tensor(0, "h", 4.0 * h)
    e.aggregateGCN(c_f)
    assert np.array_equal(e.get_tensor(1, "ah"), 4.0 * ah)
    # --- oracle on a destination-row sample (a reference-style chunk): heaviest rows included
    e.set_tensor(0, "h", h)
    e.aggregateGCN(c_f)
    deg = np.diff(g.col_ptrs)
    lo = int(np.argmax(deg))
    lo = max(0, min(lo, V - 2048))
    want = np.zeros((V, F), np.float32)
    oracle.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, h, None, low=lo, up=lo + 2048, out=want)
    assert rel_err(ah[lo:lo + 2048], want[lo:lo + 2048]) < 1e-5


def test_layer0_width_602_sample_vs_oracle(reddit, oracle):
    spec, g, e = reddit
    V, F = spec.num_vertices, spec.dims[0]
    x = synth.generate_features(V, F, 5)
    e.set_tensor(0, "x", x)
    e.aggregateGCN(e.whole_chunk(0, FORWARD))
    ah = e.get_tensor(0, "ah")
    want = np.zeros((V, F), np.float32)
    for lo in (0, 100_000, V - 1500):
        oracle.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, x, None, low=lo, up=lo + 1500, out=want)
        assert rel_err(ah[lo:lo + 1500], want[lo:lo + 1500]) < 1e-5
    # the same rows through the reference's chunk API (sub-range call): hub rows are summed by one warp
    # instead of a CTA there, so the comparison is to rounding, not to the bit
    e.aggregateGCN(Chunk(0, 0, 100_000, 101_500, 0, FORWARD, 1, True))
    assert rel_err(e.get_tensor(0, "ah")[100_000:101_500], ah[100_000:101_500]) < 1e-6
