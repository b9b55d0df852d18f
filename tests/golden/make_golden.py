"""Regenerates tests/golden/*.npz from the reference tree (run in the build container only).

    python tests/golden/make_golden.py

Sources of truth (all under /root/reference, never copied as source code):
  * miscs/dgl-non-sampling/data/raw0, raw1  -- text dumps of WeightServer::xavierInitializer
    (602x128 and 128x41, seed 8888) that the reference authors used to make DGL "computationally the
    same"; we keep raw1 whole and a row subsample of raw0.
  * miscs/check-correctness/weights-602-1000-41 -- an older weight dump: 643,000 consecutive draws of
    the same seeded generator, scaled by 1.5 (a known-answer test for the RNG stream).
  * miscs/numpy-gnn (load_data.py, layers.py, loss.py) -- the reference's dense-numpy GCN, imported and
    run on a small graph: forward tensors, soft-max and backward chain of a 2-layer model.
  * miscs/dgl-non-sampling/data/{tm,vm,sm}.pt + gendata.py -- the 66/10/24 % mask layout.
  * the reference's own loader / Matrix::dot / AdamOptimizer compiled in oracle/_ref and run on
    small seeded inputs (graph.<id>.bin images, sgemm results, Adam trajectories).
  * src/funcs/{gcn,gat}/ops/{forward,backward}_ops.cpp -- the Lambda functions' tensor ops (the
    reference's second statement of ApplyVertex / ApplyEdge), compiled into oracle/_ref and sequenced as
    funcs/gcn/main.cpp and funcs/gat/main.cpp do: soft-max, the float-wise maskout (quirk Q6), the scaled
    output gradient, tanh / tanh', weight gradients, GAT edge scores and their backward (funcs_ops.npz).
The GPU box has no /root/reference: tests read only the .npz files written here.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

from dorylus_b200 import formats  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402


def xavier_fixture():
    d = os.path.join(REF, "miscs/dgl-non-sampling/data")
    raw0 = np.loadtxt(os.path.join(d, "raw0"), dtype=np.float64)
    raw1 = np.loadtxt(os.path.join(d, "raw1"), dtype=np.float64)
    assert raw0.shape == (602, 128) and raw1.shape == (128, 41), (raw0.shape, raw1.shape)
    rows0 = np.unique(np.concatenate([np.arange(4), np.arange(0, 602, 13), [601]]))
    np.savez_compressed(os.path.join(HERE, "xavier.npz"), raw0_rows=rows0, raw0=raw0[rows0], raw1=raw1,
                        raw0_sum=raw0.sum(), raw0_abs_sum=np.abs(raw0).sum())


def weight_dump_fixture():
    """miscs/check-correctness/weights-602-1000-41: a 602x1000 and a 1000x41 matrix in the
    "Matrix Dims: (r, c)" text format (parsed by miscs/numpy-gnn/load_data.py:84-107), 6 significant
    digits.  It predates the per-matrix re-seeding: both matrices are 1.5 x consecutive draws of ONE
    std::default_random_engine(8888) + uniform_real_distribution<float>(-1, 1) stream -- a known-answer
    test for 643,000 draws of the generator WeightServer::xavierInitializer still uses.  We keep a
    subsample (flat positions in the concatenated stream + values)."""
    import re

    txt = open(os.path.join(REF, "miscs/check-correctness/weights-602-1000-41")).read()
    blocks = re.split(r"Matrix Dims: \((\d+), (\d+)\)\n", txt)
    dims, vals = [], []
    for i in range(1, len(blocks), 3):
        r, c = int(blocks[i]), int(blocks[i + 1])
        v = np.array(blocks[i + 2].split(), dtype=np.float64)
        assert v.size == r * c, (r, c, v.size)
        dims.append((r, c))
        vals.append(v)
    assert dims == [(602, 1000), (1000, 41)], dims
    stream = np.concatenate(vals)
    pos = np.unique(np.concatenate([np.arange(64), np.arange(0, stream.size, 997), np.arange(602000 - 32, 602000 + 64),
                                    np.arange(stream.size - 64, stream.size)]))
    np.savez_compressed(os.path.join(HERE, "weight_dump.npz"), dims=np.array(dims), pos=pos, vals=stream[pos],
                        total=stream.size, abs_sum=np.abs(stream).sum())


def numpy_gnn_fixture():
    """The reference's dense-numpy GCN (miscs/numpy-gnn: load_data.py builds A_hat, layers.py holds
    Aggregate / Linear / Tanh, loss.py the soft-max) is a Python reference: it is imported here and run
    on a small symmetric, duplicate-free graph (where its column-sum normalisation equals the C++
    loader's) written in the reference's own binary formats.  Stored: inputs, the
    forward tensors of a 2-layer GCN (A X W0 -> tanh -> A H W1), the soft-max, and numpy-gnn's
    backward chain (Linear / Tanh / Aggregate .backward) driven by an upstream gradient."""
    import contextlib
    import io

    sys.path.insert(0, os.path.join(REF, "miscs/numpy-gnn"))
    import layers as nl  # noqa: E402  (the reference's module)
    import load_data as nld  # noqa: E402
    import loss as nloss  # noqa: E402

    rng = np.random.default_rng(2024)
    V, dims = 90, [12, 8, 5]
    pairs = set()
    while len(pairs) < 400:
        a, b = (int(x) for x in rng.integers(0, V, 2))
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    und = np.array(sorted(pairs), dtype=np.uint32)
    src = np.concatenate([und[:, 0], und[:, 1]])
    dst = np.concatenate([und[:, 1], und[:, 0]])
    feats = (rng.random((V, dims[0]), dtype=np.float32) * 2 - 1).astype(np.float32)
    labels = rng.integers(0, dims[2], V).astype(np.uint32)
    W0 = (rng.standard_normal((dims[0], dims[1])) * 0.4).astype(np.float32)
    W1 = (rng.standard_normal((dims[1], dims[2])) * 0.4).astype(np.float32)
    upstream = (rng.standard_normal((V, dims[2])) * 1e-2).astype(np.float32)  # d handed to the backward chain
    d = tempfile.mkdtemp() + "/"
    formats.write_bsnap_edges(d + "graph.bsnap", V, src, dst)
    formats.write_features(d + "features.bsnap", feats)
    formats.write_labels(d + "labels.bsnap", labels, dims[2])
    with contextlib.redirect_stdout(io.StringIO()):
        A_hat, X, y = nld.load_data(d, "t", binary=True)
    assert A_hat.shape == (V, V) and np.array_equal(X, feats) and np.array_equal(y, labels.astype(np.int32))
    net = [nl.Aggregate("A0", A_hat), nl.Linear("W0", dims[0], dims[1], "uniform").set_W(W0.astype(np.float64)),
           nl.Tanh("t0"), nl.Aggregate("A1", A_hat), nl.Linear("W1", dims[1], dims[2], "uniform").set_W(W1.astype(np.float64))]
    acts = [X.astype(np.float64)]
    for layer in net:
        acts.append(layer.forward(acts[-1]))
    ah0, z0, h0, ah1, logits = acts[1:]
    target = np.eye(dims[2])[y]
    prob = nloss.SoftmaxCrossEntropyLoss("l").backward(logits, target) + target  # backward() returns prob - target
    g = upstream.astype(np.float64)
    grads = {}
    for layer in reversed(net):
        g = layer.backward(g)
        grads[layer.name] = g
    np.savez_compressed(os.path.join(HERE, "numpy_gnn.npz"), V=V, dims=np.array(dims), src=src, dst=dst, feats=feats,
                        labels=labels, W0=W0, W1=W1, upstream=upstream, A_hat=A_hat, ah0=ah0, z0=z0, h0=h0, ah1=ah1,
                        logits=logits, prob=prob, dW1=net[4].grad_W, grad1=grads["W1"], aTg0=grads["A1"],
                        g0=grads["t0"], dW0=net[1].grad_W)


def mask_fixture():
    import torch

    d = os.path.join(REF, "miscs/dgl-non-sampling/data")
    tm, vm, sm = (torch.load(os.path.join(d, n)).numpy() for n in ("tm.pt", "vm.pt", "sm.pt"))
    blk = 232965 // 60
    np.savez_compressed(os.path.join(HERE, "masks.npz"), vtcs=232965, parts=60, block=blk,
                        train_count=int(tm.sum()), val_count=int(vm.sum()), test_count=int(sm.sum()),
                        first_block_train=tm[:blk], first_block_val=vm[:blk])


def ref_fixture():
    ref = Ref()
    rng = np.random.default_rng(2024)
    out = {}
    # ---- loader: a 60-vertex graph with self loops, duplicates, one-directional edges, 3 parts
    V, P = 60, 3
    src = rng.integers(0, V, 420).astype(np.uint32)
    dst = rng.integers(0, V, 420).astype(np.uint32)
    src[:6] = dst[:6]  # self loops (dropped on read)
    src[6:12], dst[6:12] = src[12:18], dst[12:18]  # duplicates (kept)
    parts = rng.integers(0, P, V).astype(np.int32)
    parts[:3] = [0, 1, 2]
    out.update(ld_src=src, ld_dst=dst, ld_parts=parts, ld_V=V, ld_P=P)
    for und in (0, 1):
        d = tempfile.mkdtemp() + "/"
        formats.write_bsnap_edges(d + "graph.bsnap.edges", V, src, dst)
        formats.write_parts(d + "graph.bsnap.parts", parts)
        for p in range(P):
            f = ref.preprocess(d, p, P, bool(und))
            out["graph_u%d_p%d" % (und, p)] = np.frombuffer(open(f, "rb").read(), dtype=np.uint8)
    # ---- a single-partition graph (config 1 style) of 40 vertices
    V1 = 40
    s1 = rng.integers(0, V1, 300).astype(np.uint32)
    d1 = rng.integers(0, V1, 300).astype(np.uint32)
    dd = tempfile.mkdtemp() + "/"
    formats.write_bsnap_edges(dd + "graph.bsnap.edges", V1, s1, d1)
    formats.write_parts(dd + "graph.bsnap.parts", np.zeros(V1, np.int32))
    f = ref.preprocess(dd, 0, 1, False)
    out.update(single_src=s1, single_dst=d1, single_graph=np.frombuffer(open(f, "rb").read(), dtype=np.uint8))
    # ---- Matrix::dot, the four transpose cases
    A = rng.standard_normal((37, 19)).astype(np.float32)
    B = rng.standard_normal((19, 11)).astype(np.float32)
    Bt = np.ascontiguousarray(B.T)
    At = np.ascontiguousarray(A.T)
    out.update(dot_A=A, dot_B=B,
               dot_nn=ref.dot(A, B), dot_nt=ref.dot(A, Bt, False, True), dot_tn=ref.dot(At, B, True, False),
               dot_tt=ref.dot(At, Bt, True, True), dot_scaled=ref.dot(A, B, scale=0.5))
    # ---- AdamOptimizer: 4 synchronous "epochs" (layer 1 then layer 0, as the weight server applies them)
    dims = [6, 5, 3]
    w = [rng.standard_normal((dims[i], dims[i + 1])).astype(np.float32) for i in range(2)]
    out.update(adam_dims=np.asarray(dims), adam_w0_init=w[0].copy(), adam_w1_init=w[1].copy())
    adam = ref.adam(0.01, dims)
    grads = []
    for ep in range(4):
        g1 = rng.standard_normal(w[1].shape).astype(np.float32)
        g0 = rng.standard_normal(w[0].shape).astype(np.float32)
        grads.append((g0, g1))
        adam.update(1, w[1], g1)
        adam.update(0, w[0], g0)
        out["adam_w0_ep%d" % ep] = w[0].copy()
        out["adam_w1_ep%d" % ep] = w[1].copy()
        out["adam_g0_ep%d" % ep] = g0
        out["adam_g1_ep%d" % ep] = g1
    adam.close()
    np.savez_compressed(os.path.join(HERE, "reference_runs.npz"), **out)


def funcs_fixture():
    """Real runs of the Lambda functions' ops on seeded inputs.  V = 53 makes the maskout quirk bite in
    the middle of a row: stt = 34, (end - stt) = 19 floats = 2 rows of C = 7 plus 5 floats."""
    ref = Ref()
    rng = np.random.default_rng(77)
    out = {}
    V, Fin, Fhid, C, gV = 53, 10, 6, 7, 53  # one partition: globalVtxCnt == V
    ah0 = rng.standard_normal((V, Fin)).astype(np.float32)
    W0 = (rng.standard_normal((Fin, Fhid)) * 0.5).astype(np.float32)
    z0, h0 = ref.funcs_gcn_forward(ah0, W0)
    out.update(gcn_ah0=ah0, gcn_W0=W0, gcn_z0=z0, gcn_h0=h0)
    ah1 = rng.standard_normal((V, Fhid)).astype(np.float32)
    W1 = (rng.standard_normal((Fhid, C)) * 0.5).astype(np.float32)
    labels = rng.integers(0, C, V)
    lab = np.zeros((V, C), np.float32)
    lab[np.arange(V), labels] = 1
    scale = np.float32(gV * 0.66)  # CPU_comm.cpp:121 (float); the Lambda payload truncates it to an integer
    fin = ref.funcs_gcn_final(ah1, W1, lab, float(scale))
    out.update(gcn_ah1=ah1, gcn_W1=W1, gcn_lab=lab, gcn_gV=gV, gcn_scale=scale, gcn_pred=fin["pred"],
               gcn_masked=fin["masked"], gcn_d=fin["d"], gcn_grad1=fin["grad"], gcn_dW1=fin["dW"],
               gcn_correct_all_rows=fin["correct"], gcn_loss_all_rows=fin["loss"])
    aTg = (rng.standard_normal((V, Fhid)) * 0.1).astype(np.float32)
    rgrad, dW0 = ref.funcs_gcn_backward(ah0, z0, aTg, W0)
    out.update(gcn_aTg0=aTg, gcn_resultGrad0=rgrad, gcn_dW0=dW0)
    # ---- GAT edge ops on a ragged adjacency (empty rows included)
    Vg, Fg = 17, 5
    deg = rng.integers(0, 6, Vg)
    deg[3] = 0
    ptrs = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint64)
    z = rng.standard_normal((Vg, Fg)).astype(np.float32)
    a = rng.standard_normal((Fg, 1)).astype(np.float32)
    az, A = ref.funcs_gat_edge_forward(z, a, ptrs)
    grad = rng.standard_normal((Vg, Fg)).astype(np.float32)
    dA, dAct = ref.funcs_gat_edge_backward(grad, az, z, a, ptrs)
    ed = ref.funcs_gat_expand_dot(z, a, ptrs)
    out.update(gat_ptrs=ptrs, gat_z=z, gat_a=a, gat_az=az, gat_A=A, gat_grad=grad, gat_dA=dA, gat_dAct=dAct,
               gat_expand_dot=ed)
    np.savez_compressed(os.path.join(HERE, "funcs_ops.npz"), **out)


def input_tools_fixture():
    """inputs/graphToBinary.cpp, generateFeatues.cpp and generateLabels.cpp compiled as they are
    (oracle/build.py: build_ref_tools) and run on small inputs: the .bsnap bytes of four text edge lists
    (both --undirected settings) and the generators' output files."""
    import subprocess

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import build as ob
    from test_formats import GRAPH_TEXTS

    tools = ob.build_ref_tools()
    out = {}
    d = tempfile.mkdtemp()
    for name, text in GRAPH_TEXTS.items():
        for und in (0, 1):
            t = os.path.join(d, "%s_%d.txt" % (name, und))
            open(t, "w").write(text)
            subprocess.run([tools["graphToBinary"], "--snapfile=%s" % t, "--undirected=%d" % und, "--header=1"],
                           check=True, capture_output=True)
            out["bsnap_%s_%d" % (name, und)] = np.frombuffer(open(t + ".bsnap", "rb").read(), dtype=np.uint8)
    base = os.path.join(d, "ds")
    subprocess.run([tools["generateFeatues"], "37", "24", base], check=True, capture_output=True)
    subprocess.run([tools["generateLabels"], "37", "4", base], check=True, capture_output=True)
    out["gen_feats"] = np.frombuffer(open(base + ".feats", "rb").read(), dtype=np.uint8)
    out["gen_labels"] = np.frombuffer(open(base + ".labels", "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "input_tools.npz"), **out)


def ref_engine_fixture():
    """One synchronous GCN epoch on the reference's OWN object code (Engine::aggregateGCN, CPUComm::NNCompute:
    oracle/_ref/librefengine.so, built from gcn_ops.cpp / CPU_comm.cpp in place by oracle/build.py) on a small
    graph with the Reddit widths, Xavier weights (seed 8888): every tensor of the path, both weight updates and the
    validation statistics.  tests/test_oracle.py holds the oracle port bit-identical to it."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import random_dataset
    from oracle.pyoracle import Oracle, RefEngine

    dims = [602, 128, 41]
    ds = random_dataset(V=300, E_und=2400, dims=dims, seed=77)
    o = Oracle()
    ref = RefEngine(ds.images[0], dims)
    ref.tensor(0, "x")[:] = ds.feats
    ref.tensor(1, "lab")[:] = ds.onehot
    for l in range(2):
        ref.set_weights(l, o.xavier(dims[l], dims[l + 1]))
    ref.epoch_gcn()
    acc, loss = ref.stats()
    np.savez_compressed(os.path.join(HERE, "ref_engine.npz"), V=ds.V, dims=np.array(dims), src=ds.src, dst=ds.dst,
                        feats=ds.feats, labels=ds.labels, ah0=ref.tensor(0, "ah"), z0=ref.tensor(0, "z"), h0=ref.tensor(0, "h"),
                        ah1=ref.tensor(1, "ah"), grad1=ref.tensor(1, "grad"), aTg0=ref.tensor(0, "aTg"), dW0=ref.update(0),
                        dW1=ref.update(1), acc=acc, loss=loss)


if __name__ == "__main__":
    input_tools_fixture()
    funcs_fixture()
    xavier_fixture()
    weight_dump_fixture()
    numpy_gnn_fixture()
    ref_engine_fixture()
    mask_fixture()
    ref_fixture()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
