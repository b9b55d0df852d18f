"""tcgen05 (3xTF32) dense apply against float64 and against the fp32 CUDA-core path."""
import numpy as np
import pytest

from helpers import random_dataset, rel_err
from dorylus_b200.engine import BACKWARD, FORWARD, GCN, Engine

pytestmark = pytest.mark.gpu


def _engine(ds, tc):
    e = Engine(ds.dims, GCN)
    e.set_option("tensor_cores", 1 if tc else 0)
    e.load_partition(ds.images[0])
    e.set_tensor(len(ds.dims) - 2, "lab", ds.onehot)
    e.init_weights()
    return e


@pytest.mark.parametrize("dims,V", [([602, 128, 41], 3000), ([128, 64, 64, 25], 1111), ([96, 128, 7], 257),
                                    ([100, 64, 64, 25], 4099), ([64, 32, 5], 777), ([32, 20, 5], 130)],
                         ids=["reddit", "amazon", "odd", "amazon-100", "n32", "k32-n32"])
def test_forward_apply_tensor_core_path(dims, V):
    ds = random_dataset(V=V, E_und=4 * V, dims=dims, seed=91)
    rng = np.random.default_rng(5)
    ah = rng.standard_normal((V, dims[0])).astype(np.float32)
    out = {}
    for tc in (0, 1):
        with _engine(ds, tc) as e:
            e.set_tensor(0, "ah", ah)
            W = e.get_weights(0)
            e.applyVertexGCN(e.whole_chunk(0, FORWARD))
            out[tc] = (e.get_tensor(0, "z"), e.get_tensor(0, "h"), e.stats()["kernel_launches"])
    z64 = ah.astype(np.float64) @ W.astype(np.float64)
    err_simt, err_tc = rel_err(out[0][0], z64), rel_err(out[1][0], z64)
    print("dims", dims, "rel err vs float64: simt %.2e  tcgen05 3xTF32 %.2e" % (err_simt, err_tc))
    assert err_tc < 1e-5 and err_simt < 1e-5
    assert rel_err(out[1][1], np.tanh(z64)) < 1e-5
    assert rel_err(out[1][0], out[0][0]) < 1e-5


@pytest.mark.parametrize("dims,V", [([602, 128, 41], 5003), ([128, 64, 64, 25], 4099), ([96, 128, 7], 1025),
                                    ([602, 128, 41], 40000)],
                         ids=["reddit", "amazon", "odd", "reddit-40k"])
def test_weight_gradient_tensor_core_path(dims, V):
    """dW = AH^T . (aTg * (1 - h^2)): the contraction over the vertices on tcgen05 with MN-major
    operands (gemm_tn_tc_kernel), vertex counts that are not a multiple of the 32-vertex TMA box and
    that need several splits, against float64 and against the fp32 CUDA-core path."""
    ds = random_dataset(V=V, E_und=2 * V, dims=dims, seed=92)
    rng = np.random.default_rng(6)
    ah = rng.standard_normal((V, dims[0])).astype(np.float32)
    aTg = rng.standard_normal((V, dims[1])).astype(np.float32)
    h = np.tanh(rng.standard_normal((V, dims[1]))).astype(np.float32)
    out = {}
    for tc in (0, 1):
        with _engine(ds, tc) as e:
            e.set_tensor(0, "ah", ah)
            e.set_tensor(0, "aTg", aTg)
            e.set_tensor(0, "h", h)
            e.applyVertexGCN(e.whole_chunk(1, BACKWARD))  # NNCompute on the incremented chunk: layer 0
            out[tc] = e.get_weight_grad(0)
    g64 = aTg.astype(np.float64) * (1.0 - h.astype(np.float64) ** 2)
    dw64 = ah.astype(np.float64).T @ g64
    err_simt, err_tc = rel_err(out[0], dw64), rel_err(out[1], dw64)
    print("dims", dims, "V", V, "dW rel err vs float64: simt %.2e  tcgen05 3xTF32 %.2e" % (err_simt, err_tc))
    assert err_tc < 1e-5 and err_simt < 1e-5
    assert rel_err(out[1], out[0]) < 1e-5


def test_epochs_with_tensor_cores_match_oracle(oracle):
    from oracle.driver import OracleGCN

    ds = random_dataset(V=1500, E_und=12000, dims=[602, 128, 41], seed=93)
    orc = OracleGCN(oracle, ds.graphs, ds.dims)
    orc.load_features(ds.feats, ds.onehot)
    with _engine(ds, 1) as e:
        e.set_tensor(0, "x", ds.feats)
        for ep in range(2):
            want = orc.epoch()
            st = e.epoch()
            for l, n in ((0, "ah"), (0, "z"), (0, "h"), (1, "ah"), (1, "grad"), (0, "aTg")):
                assert rel_err(e.get_tensor(l, n), orc.saved[0][l][n]) < 1e-5, (ep, l, n)
            for l in range(2):
                assert rel_err(e.get_weight_grad(l), orc.dW[0][l]) < 1e-5
                e.set_weights(l, orc.W[l])
            assert st["acc_sum"] == want["acc"][0]


@pytest.mark.parametrize("dims,V", [([602, 128, 41], 5003), ([100, 64, 64, 25], 4099), ([16, 48, 51], 6001),
                                    ([96, 128, 64], 130), ([32, 32, 17], 127)],
                         ids=["reddit", "amazon", "friendster", "64-classes", "one-tile"])
def test_last_layer_softmax_on_tensor_cores(dims, V):
    """Last ApplyVertex (logits, soft-max, validation statistics, maskout, d, d.W^T, AH^T.d): the tcgen05
    kernel whose epilogue reads one vertex row per thread out of TMEM (gemm_tc_softmax_kernel, 1 / 2 / 4
    stages per CTA) against the separate fp32 GEMM + soft-max kernels and the fused fp32 kernel."""
    L = len(dims) - 1
    ds = random_dataset(V=V, E_und=2 * V, dims=dims, seed=94)
    rng = np.random.default_rng(7)
    ah = rng.standard_normal((V, dims[L - 1])).astype(np.float32)
    out = {}
    variants = {"separate": dict(tensor_cores=0, fuse_softmax=0), "fused-simt": dict(fuse_softmax=2),
                "tc": dict(), "tc-1": dict(tc_stages=1), "tc-2": dict(tc_stages=2), "tc-4": dict(tc_stages=4),
                "round-1 tc": dict(tc_small=0)}
    for name, opts in variants.items():
        e = Engine(ds.dims, GCN)
        for k, v in opts.items():
            e.set_option(k, v)
        e.load_partition(ds.images[0])
        e.set_tensor(L - 1, "lab", ds.onehot)
        e.init_weights()
        with e:
            e.set_tensor(L - 1, "ah", ah)
            before = e.stats()["kernel_launches"]
            e.applyVertexGCN(e.whole_chunk(L - 1, FORWARD))
            st = e.stats()
            out[name] = (e.get_tensor(L - 1, "grad"), e.get_weight_grad(L - 1), st["acc_sum"], st["loss_sum"],
                         st["kernel_launches"] - before)
    ref = out["separate"]
    assert np.abs(ref[0]).max() > 0 and np.abs(ref[1]).max() > 0
    for name, got in out.items():
        assert rel_err(got[0], ref[0]) < 1e-5, name
        assert rel_err(got[1], ref[1]) < 1e-5, name
        assert got[2] == ref[2], name
        assert abs(got[3] - ref[3]) <= 1e-5 * abs(ref[3]), name
    # the tensor-core variants really took the tcgen05 kernel: weight split + GEMM/soft-max + statistics,
    # against GEMM + soft-max + statistics (+ the split when the separate GEMM is on tensor cores too)
    print({k: v[4] for k, v in out.items()})
    nt_tc = 16 < dims[L - 1] <= 64 and 16 < dims[L] <= 128  # grad = d . W^T on the small-tile kernel: one more launch (the split)
    assert out["tc"][4] == out["round-1 tc"][4] + 1 + (1 if nt_tc else 0)
