"""Pins the CPU oracle (oracle/oracle.cpp) against the reference's golden vectors, the compiled
reference (oracle/_ref) and the dense numpy statement of the model (miscs/numpy-gnn)."""
import numpy as np
import pytest

from helpers import dense_normalized_adjacency, random_dataset, rel_err
from oracle.driver import BACKWARD, FORWARD, OracleGCN


# ---------------------------------------------------------------- golden vectors of the reference
def test_xavier_matches_reference_dumps(oracle, golden):
    """raw0 / raw1 are the reference authors' dumps of xavierInitializer (seed 8888), 8 decimals."""
    g = golden["xavier"]
    w0 = oracle.xavier(602, 128)
    w1 = oracle.xavier(128, 41)
    assert np.max(np.abs(w0[g["raw0_rows"]].astype(np.float64) - g["raw0"])) <= 6e-9
    assert np.max(np.abs(w1.astype(np.float64) - g["raw1"])) <= 6e-9
    assert abs(float(w0.astype(np.float64).sum()) - float(g["raw0_sum"])) < 1e-4


def test_rng_stream_matches_reference_weight_dump(oracle, golden):
    """miscs/check-correctness/weights-602-1000-41 (6 significant digits): 602x1000 then 1000x41
    values = 1.5 x consecutive draws of one default_random_engine(8888) / uniform(-1, 1) stream -- the
    generator xavierInitializer uses.  643,000 draws of the oracle's stream must reproduce it."""
    g = golden["weight_dump"]
    total = int(g["total"])
    assert total == 602 * 1000 + 1000 * 41
    rows = (total + 999) // 1000
    w = oracle.xavier(rows, 1000).astype(np.float64) / np.sqrt(6.0 / (rows + 1000))  # undo the Xavier scale
    stream = 1.5 * w.ravel()[:total]
    want = g["vals"]
    got = stream[g["pos"]]
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-2)) < 2e-5  # printed with 6 digits
    assert abs(np.abs(stream).sum() - float(g["abs_sum"])) < 1e-5 * float(g["abs_sum"])


def test_oracle_matches_numpy_gnn_reference_run(oracle, golden):
    """tests/golden/numpy_gnn.npz was produced by IMPORTING the reference's dense-numpy GCN
    (miscs/numpy-gnn) and running it on a small symmetric graph in the reference's file formats.
    The oracle's restatement of aggregateGCN / vtxNNForwardGCN / vtxNNBackwardGCN must reproduce its
    forward tensors, its soft-max and -- fed the same upstream gradient -- its backward chain."""
    from dorylus_b200 import engine as dengine
    from dorylus_b200 import formats

    g = golden["numpy_gnn"]
    V, dims = int(g["V"]), [int(x) for x in g["dims"]]
    image = dengine.preprocess_edges(g["src"], g["dst"], np.zeros(V, np.int32), V, 0, 1)
    pg = formats.parse_graph_bin(image)
    feats, W0, W1 = g["feats"], g["W0"], g["W1"]
    tol = 2e-6
    # the loader's arrays realise numpy-gnn's A_hat (symmetric graph: Q2 does not bite)
    A = dense_normalized_adjacency(V, g["src"], g["dst"])
    assert np.max(np.abs(A - g["A_hat"])) < 1e-12
    # forward
    ah0 = oracle.aggregate_gcn(pg.col_ptrs, pg.row_idxs, pg.fwd_vals, pg.norms, feats, None)
    assert rel_err(ah0, g["ah0"]) < tol
    z0, h0 = oracle.vtx_forward_gcn_hidden(ah0, W0)
    assert rel_err(z0, g["z0"]) < tol and rel_err(h0, g["h0"]) < tol
    ah1 = oracle.aggregate_gcn(pg.col_ptrs, pg.row_idxs, pg.fwd_vals, pg.norms, h0, None)
    assert rel_err(ah1, g["ah1"]) < tol
    onehot = formats.one_hot(g["labels"], dims[2])
    r = oracle.vtx_forward_gcn_last(ah1, W1, onehot, V)
    assert rel_err(r["pred"], g["prob"]) < tol  # un-masked soft-max of A H W1
    # backward chain from numpy-gnn's upstream gradient d: dW1 = ah1^T d, grad1 = d W1^T,
    # aTg0 = A^T grad1, g0 = aTg0 * (1 - h0^2), dW0 = ah0^T g0
    d = g["upstream"]
    dW1 = oracle.dot(ah1, d, True, False)
    grad1 = oracle.dot(d, W1, False, True)
    assert rel_err(dW1, g["dW1"]) < tol and rel_err(grad1, g["grad1"]) < tol
    aTg0 = oracle.aggregate_gcn(pg.row_ptrs, pg.col_idxs, pg.bwd_vals, pg.norms, grad1, None)
    assert rel_err(aTg0, g["aTg0"]) < tol
    dW0, _ = oracle.vtx_backward_gcn(aTg0, z0, ah0, W0, False)
    assert rel_err(dW0, g["dW0"]) < tol


def test_mask_layout_matches_gendata(golden):
    """gendata.py: per block of V/60 vertices, first int(blk*0.66) train, next int(blk*0.1) val."""
    m = golden["masks"]
    blk = int(m["block"])
    tr = int(blk * 0.66)
    va = int(blk * 0.1)
    assert m["first_block_train"][:tr].all() and not m["first_block_train"][tr:].any()
    assert m["first_block_val"][tr:tr + va].all() and not m["first_block_val"][:tr].any()
    assert int(m["train_count"]) == 60 * tr and int(m["val_count"]) == 60 * va


# ---------------------------------------------------------------- reference-generated fixtures
@pytest.mark.parametrize("case", ["nn", "nt", "tn", "tt"])
def test_dot_matches_reference_matrix_dot(oracle, golden, case):
    r = golden["reference_runs"]
    A, B = r["dot_A"], r["dot_B"]
    args = dict(nn=(A, B, False, False), nt=(A, np.ascontiguousarray(B.T), False, True),
                tn=(np.ascontiguousarray(A.T), B, True, False),
                tt=(np.ascontiguousarray(A.T), np.ascontiguousarray(B.T), True, True))[case]
    got = oracle.dot(*args)
    assert np.array_equal(got, r["dot_" + case])  # same OpenBLAS, same call => same bits
    assert rel_err(got, A.astype(np.float64) @ B.astype(np.float64)) < 1e-6


def test_dot_scale(oracle, golden):
    r = golden["reference_runs"]
    assert np.array_equal(oracle.dot(r["dot_A"], r["dot_B"], scale=0.5), r["dot_scaled"])


def test_adam_matches_reference_trajectory(oracle, golden):
    r = golden["reference_runs"]
    dims = [int(x) for x in r["adam_dims"]]
    w = [r["adam_w0_init"].copy(), r["adam_w1_init"].copy()]
    adam = oracle.adam(0.01, dims)
    for ep in range(4):
        adam.update(1, w[1], r["adam_g1_ep%d" % ep])
        adam.update(0, w[0], r["adam_g0_ep%d" % ep])
        assert np.array_equal(w[0], r["adam_w0_ep%d" % ep])
        assert np.array_equal(w[1], r["adam_w1_ep%d" % ep])
    adam.close()


def test_adam_matches_compiled_reference_live(oracle, ref):
    rng = np.random.default_rng(3)
    dims = [9, 7, 4]
    w_o = [rng.standard_normal((dims[i], dims[i + 1])).astype(np.float32) for i in range(2)]
    w_r = [w.copy() for w in w_o]
    a_o, a_r = oracle.adam(0.05, dims), ref.adam(0.05, dims)
    for _ in range(6):
        for l in (1, 0):
            g = rng.standard_normal(w_o[l].shape).astype(np.float32)
            a_o.update(l, w_o[l], g)
            a_r.update(l, w_r[l], g)
    assert np.array_equal(w_o[0], w_r[0]) and np.array_equal(w_o[1], w_r[1])


# ---------------------------------------------------------------- aggregation vs the dense statement
@pytest.mark.parametrize("P", [1, 3])
def test_aggregate_equals_dense_normalized_adjacency(oracle, P):
    ds = random_dataset(V=300, E_und=2400, dims=[24, 8, 5], P=P, seed=7)
    A = dense_normalized_adjacency(ds.V, ds.src, ds.dst)
    want_f = A @ ds.feats.astype(np.float64)
    want_b = A.T @ ds.feats.astype(np.float64)
    for g in ds.graphs:
        loc, gho = ds.feats[g.local_to_global], ds.feats[g.src_ghost_gvid]
        got = oracle.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, loc, gho)
        assert rel_err(got, want_f[g.local_to_global]) < 2e-6
        ghd = ds.feats[g.dst_ghost_gvid]
        got_b = oracle.aggregate_gcn(g.row_ptrs, g.col_idxs, g.bwd_vals, g.norms, loc, ghd)
        assert rel_err(got_b, want_b[g.local_to_global]) < 2e-6


def test_aggregate_respects_chunk_bounds(oracle):
    ds = random_dataset(V=120, E_und=600, dims=[10, 4, 3], seed=9)
    g = ds.graphs[0]
    full = oracle.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, ds.feats, None)
    part = np.full_like(full, 123.0)
    oracle.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, ds.feats, None, low=30, up=77, out=part)
    assert np.array_equal(part[30:77], full[30:77])
    assert (part[:30] == 123.0).all() and (part[77:] == 123.0).all()


# ---------------------------------------------------------------- apply vertex vs numpy
def test_vtx_forward_hidden(oracle):
    rng = np.random.default_rng(1)
    ah = rng.standard_normal((50, 12)).astype(np.float32)
    W = rng.standard_normal((12, 6)).astype(np.float32)
    z, h = oracle.vtx_forward_gcn_hidden(ah, W)
    assert rel_err(z, ah.astype(np.float64) @ W) < 1e-6
    assert rel_err(h, np.tanh(ah.astype(np.float64) @ W)) < 1e-6


def test_vtx_forward_last_replicates_maskout_quirk(oracle):
    """Q6: maskout copies (end - stt) FLOATS after row stt, not rows (CPU_comm.cpp:464-471)."""
    rng = np.random.default_rng(2)
    V, Fin, C, gV = 200, 10, 7, 200
    ah = rng.standard_normal((V, Fin)).astype(np.float32)
    W = rng.standard_normal((Fin, C)).astype(np.float32)
    lab = np.zeros((V, C), np.float32)
    lab[np.arange(V), rng.integers(0, C, V)] = 1
    r = oracle.vtx_forward_gcn_last(ah, W, lab, gV)
    z = ah.astype(np.float64) @ W
    p = np.exp(z - z.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    assert rel_err(r["pred"], p) < 1e-6
    stt = int(V * 0.66)
    flat = p.copy().reshape(-1)
    flat[stt * C: stt * C + (V - stt)] = lab.reshape(-1)[stt * C: stt * C + (V - stt)]
    d = (flat.reshape(V, C) - lab) / np.float32(gV * 0.66)
    assert rel_err(r["d"], d) < 1e-6
    assert rel_err(r["grad"], d @ W.T.astype(np.float64)) < 1e-5
    assert rel_err(r["dW"], ah.T.astype(np.float64) @ d) < 1e-5
    # validation statistics over rows [0.66 V, 0.66 V + 0.1 V)
    val = slice(stt, stt + int(V * 0.1))
    acc = float((p[val].argmax(1) == lab[val].argmax(1)).sum())
    loss = float(-np.log(p[val][np.arange(int(V * 0.1)), lab[val].argmax(1)]).sum())
    assert r["acc"] == acc and abs(r["loss"] - loss) < 1e-3


def test_vtx_backward(oracle):
    rng = np.random.default_rng(4)
    V, Fin, Fout = 64, 9, 5
    aTg, z = rng.standard_normal((V, Fout)).astype(np.float32), rng.standard_normal((V, Fout)).astype(np.float32)
    ah, W = rng.standard_normal((V, Fin)).astype(np.float32), rng.standard_normal((Fin, Fout)).astype(np.float32)
    dW, grad = oracle.vtx_backward_gcn(aTg, z, ah, W, True)
    g = aTg * (1 - np.tanh(z.astype(np.float64)) ** 2)
    assert rel_err(dW, ah.T.astype(np.float64) @ g) < 1e-5
    assert rel_err(grad, g @ W.T.astype(np.float64)) < 1e-5
    dW0, none = oracle.vtx_backward_gcn(aTg, z, ah, W, False)
    assert none is None and np.array_equal(dW0, dW)


# ---------------------------------------------------------------- the Lambda functions' own ops
def test_oracle_matches_lambda_gcn_ops_golden(oracle, golden):
    """tests/golden/funcs_ops.npz holds REAL runs of src/funcs/gcn/ops (softmax, maskout, tanh,
    tanhDerivative) sequenced as funcs/gcn/main.cpp does -- the reference's second statement of
    ApplyVertex.  The oracle's restatement of CPUComm::vtxNN{Forward,Backward}GCN must agree with it,
    including the float-wise maskout (quirk Q6) and the 1/(V * 0.66) scale."""
    g = golden["funcs_ops"]
    z0, h0 = oracle.vtx_forward_gcn_hidden(g["gcn_ah0"], g["gcn_W0"])
    assert rel_err(z0, g["gcn_z0"]) < 1e-6 and rel_err(h0, g["gcn_h0"]) < 1e-6
    r = oracle.vtx_forward_gcn_last(g["gcn_ah1"], g["gcn_W1"], g["gcn_lab"], int(g["gcn_gV"]))
    assert rel_err(r["pred"], g["gcn_pred"]) < 1e-6
    assert rel_err(r["d"], g["gcn_d"]) < 1e-6
    assert rel_err(r["grad"], g["gcn_grad1"]) < 1e-6
    assert rel_err(r["dW"], g["gcn_dW1"]) < 1e-6
    dW0, grad0 = oracle.vtx_backward_gcn(g["gcn_aTg0"], g["gcn_z0"], g["gcn_ah0"], g["gcn_W0"], True)
    assert rel_err(dW0, g["gcn_dW0"]) < 1e-6
    assert rel_err(grad0, g["gcn_resultGrad0"]) < 1e-6


def test_lambda_maskout_quirk_shape(golden):
    """What the reference's maskout() really overwrites: (end - stt) FLOATS from row stt on, i.e. two
    rows and five floats of a third at V = 53, C = 7 -- not the 19 non-training rows."""
    g = golden["funcs_ops"]
    pred, masked, lab = g["gcn_pred"], g["gcn_masked"], g["gcn_lab"]
    V, C = pred.shape
    stt = int(V * 0.66)
    want = pred.copy().reshape(-1)
    want[stt * C: stt * C + (V - stt)] = lab.reshape(-1)[stt * C: stt * C + (V - stt)]
    assert np.array_equal(masked.reshape(-1), want)
    assert (V - stt) % C != 0  # the fixture ends mid-row on purpose
    d = (masked - lab) / g["gcn_scale"]
    assert rel_err(g["gcn_d"], d) < 1e-6
    # checkAccuracy / checkLoss (forward_ops.cpp) run over all rows
    assert int(g["gcn_correct_all_rows"]) == int((pred.argmax(1) == lab.argmax(1)).sum())
    loss = float(-np.log(pred[np.arange(V), lab.argmax(1)].astype(np.float64)).sum())
    assert abs(float(g["gcn_loss_all_rows"]) - loss) < 1e-3


def test_oracle_matches_lambda_gat_ops_golden(oracle, golden):
    """src/funcs/gat/ops: edgeMatMul + leakyReLU (edge forward), leakyReLUDerivative +
    expandHadamardMul + dAct . a (edge backward), expandDot -- on a ragged adjacency with empty rows."""
    g = golden["funcs_ops"]
    az, A = oracle.edg_forward_gat(g["gat_z"], g["gat_a"], g["gat_ptrs"])
    assert rel_err(az, g["gat_az"]) < 1e-6 and rel_err(A, g["gat_A"]) < 1e-6
    assert rel_err(az, g["gat_expand_dot"]) < 1e-6  # expandDot == edgeMatMul (quirk Q8: one-sided score)
    dA, _da = oracle.edg_backward_gat(g["gat_grad"], g["gat_az"], g["gat_z"], g["gat_a"], g["gat_ptrs"])
    assert rel_err(dA, g["gat_dA"]) < 1e-6
    # dAct is the E x F' tensor the reference materialises: grad[dst(e)] * leaky'(az[e])
    ptrs = g["gat_ptrs"].astype(np.int64)
    dst = np.repeat(np.arange(len(ptrs) - 1), np.diff(ptrs))
    want = g["gat_grad"][dst] * np.where(g["gat_az"] > 0, 1.0, 0.01)[:, None].astype(np.float32)
    assert rel_err(g["gat_dAct"], want) < 1e-6


@pytest.mark.parametrize("V,Fin,Fout,C", [(200, 33, 17, 5), (1, 4, 3, 2), (97, 602, 128, 41)])
def test_oracle_matches_lambda_ops_live(oracle, ref, V, Fin, Fout, C):
    """Same comparison against the compiled reference ops on fresh inputs (needs oracle/_ref)."""
    rng = np.random.default_rng(V)
    ah = rng.standard_normal((V, Fin)).astype(np.float32)
    W = (rng.standard_normal((Fin, Fout)) * 0.3).astype(np.float32)
    z, h = ref.funcs_gcn_forward(ah, W)
    oz, oh = oracle.vtx_forward_gcn_hidden(ah, W)
    assert rel_err(oz, z) < 1e-6 and rel_err(oh, h) < 1e-6
    aTg = rng.standard_normal((V, Fout)).astype(np.float32)
    rgrad, rdW = ref.funcs_gcn_backward(ah, z, aTg, W)
    odW, ograd = oracle.vtx_backward_gcn(aTg, z, ah, W, True)
    assert rel_err(odW, rdW) < 1e-6 and rel_err(ograd, rgrad) < 1e-6
    W1 = (rng.standard_normal((Fout, C)) * 0.3).astype(np.float32)
    lab = np.zeros((V, C), np.float32)
    lab[np.arange(V), rng.integers(0, C, V)] = 1
    fin = ref.funcs_gcn_final(h, W1, lab, float(np.float32(V * 0.66)))
    o = oracle.vtx_forward_gcn_last(h, W1, lab, V)
    for k in ("pred", "d", "grad", "dW"):
        assert rel_err(o[k], fin[k]) < 1e-6, k


# ---------------------------------------------------------------- whole epoch: partitions agree
def test_epoch_partitioned_equals_single_partition(oracle):
    """The per-partition state machine with ghost exchange reproduces the 1-partition run
    (multi-GPU == single-GPU equivalence, stated on the oracle).  The loss mask is per partition
    (quirk Q5: first 66 % of EACH partition's local order), so gradients are compared from a common
    grad[1] rather than through the loss."""
    dims = [20, 8, 5]
    one = random_dataset(V=240, E_und=1500, dims=dims, P=1, seed=11)
    many = random_dataset(V=240, E_und=1500, dims=dims, P=3, seed=11)
    r1, r3 = OracleGCN(oracle, one.graphs, dims), OracleGCN(oracle, many.graphs, dims)
    r1.load_features(one.feats, one.onehot)
    r3.load_features(many.feats, many.onehot)
    r1.epoch()
    for l in range(2):
        for p in range(3):
            r3.aggregate(p, l, FORWARD)
            r3.apply_vertex_forward(p, l)
        if l == 0:
            r3.scatter(1, FORWARD)
    for name, layer in (("ah", 0), ("z", 0), ("h", 0), ("ah", 1)):
        full = r1.saved[0][layer][name]
        for p, g in enumerate(many.graphs):
            assert rel_err(r3.saved[p][layer][name], full[g.local_to_global]) < 1e-5, (name, layer, p)
    for p, g in enumerate(many.graphs):
        r3.saved[p][1]["grad"][:] = r1.saved[0][1]["grad"][g.local_to_global]
    r3.scatter(1, BACKWARD)
    for p, g in enumerate(many.graphs):
        r3.aggregate(p, 1, BACKWARD)
        assert rel_err(r3.saved[p][0]["aTg"], r1.saved[0][0]["aTg"][g.local_to_global]) < 1e-5
    # weight gradients sum over partitions to the single-partition gradient (same g, same ah)
    tot = None
    for p in range(3):
        r3.apply_vertex_backward(p, 0)
        tot = r3.dW[p][0] if tot is None else tot + r3.dW[p][0]
    assert rel_err(tot, r1.dW[0][0]) < 1e-5


def test_chunk_priority_equals_reference_operator_less(ref, tmp_path):
    """The chunk queues of host/saga_pipeline.hpp are ordered by a restatement of Chunk::operator<
    (common/utils.hpp:76-89); here both are called on 20,000 random chunk pairs drawn from small field
    ranges (so that ties on every prefix of the comparison chain occur), and the Python mirror's
    isFirstLayer / isLastLayer are compared with the reference's."""
    import ctypes as C
    import os
    import subprocess

    from dorylus_b200.engine import Chunk

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libsaga_bindings.so")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", os.path.join(root, "tests", "saga_bindings.cpp"), "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    saga = C.CDLL(so)
    rng = np.random.default_rng(5)
    u32p = C.POINTER(C.c_uint)
    n_true = 0
    for _ in range(20000):
        a = rng.integers(0, 3, 8).astype(np.uint32)
        b = rng.integers(0, 3, 8).astype(np.uint32)
        for f in (a, b):
            f[5] %= 2
            f[7] %= 2
        if rng.random() < 0.3:
            k = int(rng.integers(1, 8))
            b[:k] = a[:k]  # force ties on a prefix of the fields
        want = ref.lib.ref_chunk_less(a.ctypes.data_as(u32p), b.ctypes.data_as(u32p))
        got = saga.saga_chunk_less(a.ctypes.data_as(u32p), b.ctypes.data_as(u32p))
        assert want == got, (a, b)
        n_true += want
        flags = ref.lib.ref_chunk_flags(a.ctypes.data_as(u32p))
        c = Chunk(*(int(x) for x in a[:7]), vertex=bool(a[7]))
        assert (1 if c.isFirstLayer() else 0) | (2 if c.isLastLayer() else 0) == flags
    assert 2000 < n_true < 18000


def test_sync_weight_update_matches_reference_weight_tensor(oracle, ref):
    """The weight server's synchronous update, run through the reference's own WeightTensor +
    AdamOptimizer (weighttensor.cpp compiled in oracle/_ref): gradients of 4 partitions arrive as
    "local" updates (one weight server) or as 2 local updates + 1 relayed "ghost" sum (two weight
    servers); nothing is applied until every expected update is in; then the summed gradient steps Adam.
    The oracle's apply_updates (sum over partitions in order, then Adam) must give the same weights."""
    import ctypes as C

    L = ref.lib
    L.ref_wt_create.restype = C.c_void_p
    L.ref_adam_create.restype = C.c_void_p
    f32p = C.POINTER(C.c_float)
    rng = np.random.default_rng(8)
    dims = [6, 5, 3]
    for local_tot, ghost_tot in ((4, 0), (2, 1)):
        W = [rng.standard_normal((dims[i], dims[i + 1])).astype(np.float32) for i in range(2)]
        mine = [w.copy() for w in W]
        d = np.asarray(dims, np.uint32)
        radam = C.c_void_p(L.ref_adam_create(C.c_float(0.01), d.ctypes.data_as(C.POINTER(C.c_uint32)), 3))
        oadam = oracle.adam(0.01, dims)
        wts = [C.c_void_p(L.ref_wt_create(W[l].ctypes.data_as(f32p), dims[l], dims[l + 1], local_tot, ghost_tot)) for l in range(2)]
        for ep in range(3):
            for l in (1, 0):  # the order the updates reach the weight server
                grads = [rng.standard_normal(W[l].shape).astype(np.float32) for _ in range(4)]
                out = np.empty_like(W[l])
                n = grads[0].size
                for k in range(local_tot):
                    # not ready yet: nothing is applied, the weights stay
                    assert L.ref_wt_try_apply(wts[l], radam, l, out.ctypes.data_as(f32p)) == 0
                    assert np.array_equal(out, mine[l])
                    L.ref_wt_local_update(wts[l], grads[k].ctypes.data_as(f32p), n)
                if ghost_tot:
                    relayed = grads[2] + grads[3]  # what the other weight server summed and sent over
                    assert L.ref_wt_try_apply(wts[l], radam, l, out.ctypes.data_as(f32p)) == 0
                    L.ref_wt_ghost_update(wts[l], relayed.ctypes.data_as(f32p), n)
                assert L.ref_wt_try_apply(wts[l], radam, l, out.ctypes.data_as(f32p)) == 1
                # the oracle's statement (oracle/driver.py: apply_updates)
                total = grads[0].copy()
                if ghost_tot:
                    total += grads[1]
                    total += grads[2] + grads[3]
                else:
                    for k in range(1, 4):
                        total += grads[k]
                oadam.update(l, mine[l], total)
                assert np.array_equal(out, mine[l]), (local_tot, ghost_tot, ep, l)
        for h in wts:
            L.ref_wt_destroy(h)
        L.ref_adam_destroy(radam)
        oadam.close()


# ------------------------------------------------------------------------------------------------
# The oracle port against the reference's OWN object code for the hot path: Engine::aggregateGCN
# (gcn_ops.cpp:130-191) and CPUComm::NNCompute -> vtxNNForwardGCN / vtxNNBackwardGCN with every helper
# (CPU_comm.cpp:98-159, 265-297, 424-471), compiled in place into oracle/_ref/librefengine.so
# (oracle/ref_engine.cpp, oracle/build.py: build_ref_engine).  This pins what SURVEY.md 8c had to leave
# "pinned by our oracle only": the aggregation's summation order, getTrainStat and the loss scale.
def _ref_engine_case(oracle, V, E_und, dims, seed, epochs):
    import numpy as np
    from helpers import random_dataset
    from oracle.driver import OracleGCN
    from oracle.pyoracle import RefEngine

    ds = random_dataset(V=V, E_und=E_und, dims=dims, seed=seed)
    L = len(dims) - 1
    orc = OracleGCN(oracle, ds.graphs, dims)
    orc.load_features(ds.feats, ds.onehot)
    ref = RefEngine(ds.images[0], dims)
    ref.tensor(0, "x")[:] = ds.feats
    ref.tensor(L - 1, "lab")[:] = ds.onehot
    out = []
    for ep in range(epochs):
        for l in range(L):
            ref.set_weights(l, orc.W[l])  # this epoch's weights, as the weight server would hand them out
        want = orc.epoch()               # (steps Adam at its end)
        ref.epoch_gcn()
        t = orc.saved[0]
        got = {}
        for l in range(L):
            got["ah%d" % l] = (ref.tensor(l, "ah").copy(), t[l]["ah"].copy())
            if l < L - 1:
                got["z%d" % l] = (ref.tensor(l, "z").copy(), t[l]["z"].copy())
                got["h%d" % l] = (ref.tensor(l, "h").copy(), t[l]["h"].copy())
                got["aTg%d" % l] = (ref.tensor(l, "aTg").copy(), t[l]["aTg"].copy())
            if l > 0:
                got["grad%d" % l] = (ref.tensor(l, "grad").copy(), t[l]["grad"].copy())
            got["dW%d" % l] = (ref.update(l), orc.dW[0][l].copy())
        out.append((got, ref.stats(), (want["acc"][0], want["loss"][0])))
    return out


@pytest.mark.parametrize("shape", [dict(V=600, E_und=7200, dims=[602, 128, 41]), dict(V=2708, E_und=5278, dims=[1433, 16, 7]),
                                   dict(V=2000, E_und=9000, dims=[100, 64, 64, 25]), dict(V=1800, E_und=6000, dims=[16, 48, 51])],
                         ids=lambda s: "x".join(map(str, s["dims"])))
def test_oracle_equals_reference_engine_object_code(oracle, shape):
    from oracle.pyoracle import RefEngine

    if not RefEngine.available():
        pytest.skip("oracle/_ref/librefengine.so not available (no /root/reference and no prebuilt library)")
    for got, stats, want_stats in _ref_engine_case(oracle, seed=21, epochs=2, **shape):
        for name, (a, b) in got.items():
            assert np.array_equal(a, b), name  # the port is bit-identical to the reference's own code
        assert stats == want_stats


def test_reference_engine_honours_chunk_bounds(oracle):
    """Engine::aggregateGCN on a sub-range chunk [lowBound, upBound) (Lambda-style chunking) against the port."""
    from helpers import random_dataset
    from oracle.pyoracle import RefEngine

    if not RefEngine.available():
        pytest.skip("oracle/_ref/librefengine.so not available")
    ds = random_dataset(V=500, E_und=4000, dims=[24, 8, 5], seed=9)
    g = ds.graphs[0]
    ref = RefEngine(ds.images[0], ds.dims)
    ref.tensor(0, "x")[:] = ds.feats
    ref.tensor(0, "ah")[:] = -7.0
    ref.aggregate(0, 0, 100, 333)
    want = np.full((500, 24), -7.0, np.float32)
    oracle.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, ds.feats, None, 100, 333, out=want)
    assert np.array_equal(ref.tensor(0, "ah"), want)


def test_oracle_equals_reference_engine_golden(oracle, golden):
    """The same comparison against tests/golden/ref_engine.npz -- outputs of the reference's own object code
    recorded by tests/golden/make_golden.py, for boxes where neither /root/reference nor the prebuilt library is."""
    from oracle.driver import OracleGCN
    from dorylus_b200 import engine as dengine
    from dorylus_b200 import formats

    g = golden["ref_engine"]
    V, dims = int(g["V"]), [int(x) for x in g["dims"]]
    image = dengine.preprocess_edges(g["src"], g["dst"], np.zeros(V, np.int32), V, 0, 1)
    orc = OracleGCN(oracle, [formats.parse_graph_bin(image)], dims)
    orc.load_features(g["feats"], formats.one_hot(g["labels"], dims[-1]))
    want = orc.epoch()
    t = orc.saved[0]
    for name in ("ah0", "z0", "h0", "ah1", "grad1", "aTg0", "dW0", "dW1"):
        mine = orc.dW[0][int(name[-1])] if name.startswith("dW") else t[int(name[-1])][name[:-1]]
        assert np.array_equal(mine, g[name]), name
    assert (want["acc"][0], want["loss"][0]) == (float(g["acc"]), float(g["loss"]))


@pytest.mark.parametrize("shape", [dict(V=300, E_und=3000, dims=[24, 12, 5]), dict(V=600, E_und=14000, dims=[602, 128, 41])],
                         ids=lambda s: "x".join(map(str, s["dims"])))
def test_oracle_gat_equals_reference_engine_object_code(oracle, shape):
    """GAT (BASELINE configs[2]): Engine::aggregateGAT / predictGAT / applyVertexGAT / applyEdgeGAT of gat_ops.cpp and
    CPUComm's vtxNN*GAT / edgNN*GAT, the reference's own object code, against the port -- bit for bit on every
    tensor of both layers, forward and backward, with the reference's quirks in force: predictGAT reads "az" (Q9),
    "A" aliases forwardAdj.values (Q12).  Two things the reference leaves undefined are excluded and said so:
    * "da": CPUComm::reduce() accumulates into memory operator new[] left uninitialised (Q11, CPU_comm.cpp:366-382);
    * predictGAT's `#pragma omp parallel` WITHOUT `for` (gat_ops.cpp:258-263) makes every OpenMP thread subtract the
      labels: with T threads grad = softmax - (up to T) * lab, racily.  The library runs it with one thread here;
      the port (and the engine) implement softmax - lab."""
    from helpers import random_dataset
    from oracle.driver import OracleGAT
    from oracle.pyoracle import RefEngine

    if not RefEngine.available():
        pytest.skip("oracle/_ref/librefengine.so not available")
    dims = shape["dims"]
    ds = random_dataset(seed=71, **shape)  # E >= V * C: the "az" quirk reads V x C floats of the E x 1 score vector
    orc = OracleGAT(oracle, ds.graphs, dims, predict_from="az")
    orc.load_features(ds.feats, ds.onehot)
    orc.epoch()
    ref = RefEngine(ds.images[0], dims, gat=True)
    ref.set_threads(1)
    ref.tensor(0, "h")[:] = ds.feats
    ref.tensor(1, "lab")[:] = ds.onehot
    for l in range(2):
        ref.set_weights(l, orc.W[l])
        ref.set_a(l, orc.a[l])
    for l in range(2):
        ref.gat_forward(l)
    for l in (1, 0):
        ref.gat_backward(l)
    t = orc.saved[0]
    for l in range(2):
        for n in ("z", "ah", "grad", "aTg"):
            assert np.array_equal(ref.tensor(l, n), t[l][n]), (l, n)
        for n in ("az", "dA"):
            assert np.array_equal(ref.tensor(l, n).reshape(-1), t[l][n]), (l, n)
        assert np.array_equal(ref.update(l), orc.dW[0][l]), (l, "dW")
    assert np.array_equal(ref.tensor(1, "A").reshape(-1), orc.A[0])
