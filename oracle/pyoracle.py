"""ctypes bindings of the CPU oracle (TEST INFRASTRUCTURE ONLY).

``Oracle`` wraps oracle/_build/liboracle.so (our restatement, oracle/oracle.cpp) and ``Ref`` wraps
oracle/_ref/libdoryref.so (the reference's own translation units).  Only tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def _f(a):
    return a.ctypes.data_as(_f32p)


def _u(a):
    return a.ctypes.data_as(_u32p)


def _q(a):
    return a.ctypes.data_as(_u64p)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Oracle:
    def __init__(self, build: bool = True):
        path = os.path.join(_build.BUILD_DIR, "liboracle.so")
        if build or not os.path.exists(path):
            path = _build.build_oracle()
        self.lib = L = C.CDLL(path)
        L.orc_build_edge_table.restype = C.c_void_p
        L.orc_adam_create.restype = C.c_void_p
        L.orc_adam_lr_t.restype = C.c_float
        L.orc_get_threads.restype = C.c_int

    # ------------------------------------------------------------------ threading
    def set_threads(self, n: int):
        self.lib.orc_set_threads(int(n))

    def threads(self) -> int:
        return int(self.lib.orc_get_threads())

    # ------------------------------------------------------------------ aggregation
    def edge_table(self, ptrs, idxs, local, ghost, low=0, up=None):
        """Engine::srcVFeats2eFeats / dstVFeats2eFeats.  Keeps `local`/`ghost` alive via the handle."""
        V, F = local.shape
        if ghost is None or ghost.size == 0:
            ghost = np.zeros((1, F), np.float32)
        up = V if up is None else up
        h = self.lib.orc_build_edge_table(_q(ptrs), _u(idxs), C.c_uint(V), _f(local), _f(ghost), C.c_uint(F),
                                          C.c_uint(low), C.c_uint(up))
        return (C.c_void_p(h), local, ghost)

    def free_edge_table(self, t):
        self.lib.orc_free_edge_table(t[0])

    def aggregate_gcn(self, ptrs, idxs, vals, norms, local, ghost, low=0, up=None, out=None, table=None):
        """Engine::aggregateGCN for one direction; `local`/`ghost` are the source tensors."""
        local = _c32(local)
        ghost = _c32(ghost) if ghost is not None else None
        V, F = local.shape
        up = V if up is None else up
        if out is None:
            out = np.zeros((V, F), np.float32)
        own = table is None
        if own:
            table = self.edge_table(ptrs, idxs, local, ghost)
        self.lib.orc_aggregate_gcn(_q(ptrs), _f(vals), _f(norms), _f(table[1]), table[0], C.c_uint(F),
                                   C.c_uint(low), C.c_uint(up), _f(out))
        if own:
            self.free_edge_table(table)
        return out

    def aggregate_gat_fwd(self, col_ptrs, row_idxs, A, z, z_ghost, low=0, up=None, out=None):
        z = _c32(z)
        V, F = z.shape
        up = V if up is None else up
        if out is None:
            out = np.zeros((V, F), np.float32)
        t = self.edge_table(col_ptrs, row_idxs, z, _c32(z_ghost) if z_ghost is not None else None)
        self.lib.orc_aggregate_gat_fwd(_q(col_ptrs), _f(_c32(A)), _f(t[1]), t[0], C.c_uint(F),
                                       C.c_uint(low), C.c_uint(up), _f(out))
        self.free_edge_table(t)
        return out

    def aggregate_gat_bwd(self, col_ptrs, row_idxs, dA, z, z_ghost, row_ptrs, col_idxs, bvals, grad,
                          grad_ghost, low=0, up=None, out=None):
        z = _c32(z)
        grad = _c32(grad)
        V, F = z.shape
        up = V if up is None else up
        if out is None:
            out = np.zeros((V, F), np.float32)
        tf = self.edge_table(col_ptrs, row_idxs, z, _c32(z_ghost) if z_ghost is not None else None)
        tb = self.edge_table(row_ptrs, col_idxs, grad, _c32(grad_ghost) if grad_ghost is not None else None)
        self.lib.orc_aggregate_gat_bwd(_q(col_ptrs), _f(_c32(dA)), tf[0], _q(row_ptrs), _f(bvals), tb[0],
                                       C.c_uint(F), C.c_uint(low), C.c_uint(up), _f(out))
        self.free_edge_table(tf)
        self.free_edge_table(tb)
        return out

    def predict_gat(self, logits, labels):
        logits, labels = _c32(logits), _c32(labels)
        out = np.empty_like(labels)
        self.lib.orc_predict_gat(_f(logits), _f(labels), C.c_uint(labels.shape[0]), C.c_uint(labels.shape[1]), _f(out))
        return out

    def softmax(self, x):
        x = _c32(x)
        out = np.empty_like(x)
        self.lib.orc_softmax(_f(x), C.c_uint(x.shape[0]), C.c_uint(x.shape[1]), _f(out))
        return out

    # ------------------------------------------------------------------ dense
    def dot(self, A, B, tA=False, tB=False, scale=1.0):
        A, B = _c32(A), _c32(B)
        m = A.shape[1] if tA else A.shape[0]
        n = B.shape[0] if tB else B.shape[1]
        out = np.empty((m, n), np.float32)
        self.lib.orc_matrix_dot(_f(A), C.c_uint(A.shape[0]), C.c_uint(A.shape[1]), _f(B), C.c_uint(B.shape[0]),
                                C.c_uint(B.shape[1]), int(tA), int(tB), C.c_float(scale), _f(out))
        return out

    def vtx_forward_gcn_hidden(self, ah, W):
        ah, W = _c32(ah), _c32(W)
        V, Fin = ah.shape
        Fout = W.shape[1]
        z = np.empty((V, Fout), np.float32)
        h = np.empty((V, Fout), np.float32)
        self.lib.orc_vtx_forward_gcn_hidden(_f(ah), _f(W), C.c_uint(V), C.c_uint(Fin), C.c_uint(Fout), _f(z), _f(h))
        return z, h

    def vtx_forward_gcn_last(self, ah, W, lab, global_vtx_cnt):
        ah, W, lab = _c32(ah), _c32(W), _c32(lab)
        V, Fin = ah.shape
        Cc = W.shape[1]
        pred = np.empty((V, Cc), np.float32)
        d = np.empty((V, Cc), np.float32)
        grad = np.empty((V, Fin), np.float32)
        dW = np.empty((Fin, Cc), np.float32)
        acc, loss = C.c_float(), C.c_float()
        self.lib.orc_vtx_forward_gcn_last(_f(ah), _f(W), _f(lab), C.c_uint(V), C.c_uint(Fin), C.c_uint(Cc),
                                          C.c_uint(global_vtx_cnt), _f(pred), C.byref(acc), C.byref(loss),
                                          _f(grad), _f(dW), _f(d))
        return dict(pred=pred, acc=acc.value, loss=loss.value, grad=grad, dW=dW, d=d)

    def vtx_backward_gcn(self, aTg, z, ah, W, layer_nonzero: bool):
        aTg, z, ah, W = _c32(aTg), _c32(z), _c32(ah), _c32(W)
        V, Fin = ah.shape
        Fout = W.shape[1]
        dW = np.empty((Fin, Fout), np.float32)
        grad = np.empty((V, Fin), np.float32) if layer_nonzero else np.zeros((1, 1), np.float32)
        self.lib.orc_vtx_backward_gcn(_f(aTg), _f(z), _f(ah), _f(W), C.c_uint(V), C.c_uint(Fin), C.c_uint(Fout),
                                      int(layer_nonzero), _f(dW), _f(grad))
        return dW, (grad if layer_nonzero else None)

    def vtx_forward_gat(self, feats, W):
        feats, W = _c32(feats), _c32(W)
        z = np.empty((feats.shape[0], W.shape[1]), np.float32)
        self.lib.orc_vtx_forward_gat(_f(feats), _f(W), C.c_uint(feats.shape[0]), C.c_uint(feats.shape[1]),
                                     C.c_uint(W.shape[1]), _f(z))
        return z

    def vtx_backward_gat(self, h, aTg, W, layer_nonzero: bool):
        h, aTg, W = _c32(h), _c32(aTg), _c32(W)
        V, Fin = h.shape
        Fout = W.shape[1]
        dW = np.empty((Fin, Fout), np.float32)
        grad = np.empty((V, Fin), np.float32) if layer_nonzero else np.zeros((1, 1), np.float32)
        self.lib.orc_vtx_backward_gat(_f(h), _f(aTg), _f(W), C.c_uint(V), C.c_uint(Fin), C.c_uint(Fout),
                                      int(layer_nonzero), _f(dW), _f(grad))
        return dW, (grad if layer_nonzero else None)

    def edg_forward_gat(self, z, a, col_ptrs):
        z, a = _c32(z), _c32(a).reshape(-1)
        V, F = z.shape
        nnz = int(col_ptrs[V])
        az = np.zeros(max(nnz, 1), np.float32)
        A = np.zeros(max(nnz, 1), np.float32)
        self.lib.orc_edg_forward_gat(_f(z), _f(a), _q(col_ptrs), C.c_uint(V), C.c_uint(F), _f(az), _f(A))
        return az[:nnz], A[:nnz]

    def edg_backward_gat(self, grad, az, z, a, col_ptrs):
        grad, az, z, a = _c32(grad), _c32(az), _c32(z), _c32(a).reshape(-1)
        V, F = z.shape
        nnz = int(col_ptrs[V])
        dA = np.zeros(max(nnz, 1), np.float32)
        da = np.zeros(F, np.float32)
        self.lib.orc_edg_backward_gat(_f(grad), _f(az), _f(z), _f(a), _q(col_ptrs), C.c_uint(V), C.c_uint(F),
                                      _f(dA), _f(da))
        return dA[:nnz], da

    # ------------------------------------------------------------------ weights
    def xavier(self, d1, d2):
        out = np.empty((d1, d2), np.float32)
        self.lib.orc_xavier(C.c_uint(d1), C.c_uint(d2), _f(out))
        return out

    def kaiming(self, d1, d2):
        out = np.empty((d1, d2), np.float32)
        self.lib.orc_kaiming(C.c_uint(d1), C.c_uint(d2), _f(out))
        return out

    def adam(self, lr, dims):
        return _Adam(self.lib, "orc", lr, dims)


class _Adam:
    def __init__(self, lib, prefix, lr, dims):
        self.lib, self.p = lib, prefix
        d = np.asarray(dims, dtype=np.uint32)
        getattr(lib, prefix + "_adam_create").restype = C.c_void_p
        self.h = C.c_void_p(getattr(lib, prefix + "_adam_create")(C.c_float(lr), _u(d), C.c_uint(d.size)))

    def update(self, layer, weight, grad):
        """In-place on `weight` (contiguous fp32)."""
        assert weight.dtype == np.float32 and weight.flags.c_contiguous
        grad = _c32(grad).copy()
        getattr(self.lib, self.p + "_adam_update")(self.h, C.c_uint(layer), _f(weight), _f(grad))

    def close(self):
        if self.h:
            getattr(self.lib, self.p + "_adam_destroy")(self.h)
            self.h = None


class Ref:
    """The reference's own compiled loader / Matrix / Adam (oracle/_ref/libdoryref.so)."""

    def __init__(self):
        path = _build.build_ref()
        if path is None or not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libdoryref.so unavailable (no /root/reference, no prebuilt)")
        self.lib = L = C.CDLL(path)
        L.ref_graph_open.restype = C.c_void_p
        L.ref_graph_sendlist.restype = C.c_uint

    @staticmethod
    def available() -> bool:
        try:
            p = _build.build_ref()
            return p is not None and os.path.exists(p)
        except Exception:
            return False

    def preprocess(self, dataset_dir: str, node_id: int, num_nodes: int, undirected: bool = False):
        d = dataset_dir if dataset_dir.endswith("/") else dataset_dir + "/"
        self.lib.ref_preprocess(d.encode(), C.c_uint(node_id), C.c_uint(num_nodes), int(undirected))
        return d + "graph.%d.bin" % node_id

    def load_graph(self, path: str) -> dict:
        h = C.c_void_p(self.lib.ref_graph_open(path.encode()))
        cnt = np.zeros(9, np.uint64)
        self.lib.ref_graph_counts(h, _q(cnt))
        V, gV, sg, dg, ine, oute, ge, nf, nb = (int(x) for x in cnt)
        out = dict(local_vtx_cnt=V, global_vtx_cnt=gV, src_ghost_cnt=sg, dst_ghost_cnt=dg,
                   local_in_edge_cnt=ine, local_out_edge_cnt=oute, global_edge_cnt=ge)
        a = dict(local_to_global=np.zeros(V, np.uint32), norms=np.zeros(V, np.float32),
                 col_ptrs=np.zeros(V + 1, np.uint64), row_idxs=np.zeros(nf, np.uint32),
                 fwd_vals=np.zeros(nf, np.float32), row_ptrs=np.zeros(V + 1, np.uint64),
                 col_idxs=np.zeros(nb, np.uint32), bwd_vals=np.zeros(nb, np.float32))
        self.lib.ref_graph_arrays(h, _u(a["local_to_global"]), _f(a["norms"]), _q(a["col_ptrs"]),
                                  _u(a["row_idxs"]), _f(a["fwd_vals"]), _q(a["row_ptrs"]),
                                  _u(a["col_idxs"]), _f(a["bwd_vals"]))
        out.update(a)
        for which, name, n in ((0, "src_ghost", sg), (1, "dst_ghost", dg)):
            g = np.zeros(max(n, 1), np.uint32)
            l = np.zeros(max(n, 1), np.uint32)
            self.lib.ref_graph_ghosts(h, which, _u(g), _u(l))
            out[name + "_gvid"], out[name + "_lvid"] = g[:n], l[:n]
        for which, name in ((0, "fwd_send"), (1, "bwd_send")):
            lists = []
            peer = 0
            while True:
                n = self.lib.ref_graph_sendlist(h, which, C.c_uint(peer), None)
                # lists beyond numNodes return 0; stop at a generous bound
                if peer >= 64:
                    break
                buf = np.zeros(max(n, 1), np.uint32)
                if n:
                    self.lib.ref_graph_sendlist(h, which, C.c_uint(peer), _u(buf))
                lists.append(buf[:n])
                peer += 1
            out[name] = lists
        self.lib.ref_graph_close(h)
        return out

    def dot(self, A, B, tA=False, tB=False, scale=1.0):
        A, B = _c32(A), _c32(B)
        m = A.shape[1] if tA else A.shape[0]
        n = B.shape[0] if tB else B.shape[1]
        out = np.empty((m, n), np.float32)
        self.lib.ref_matrix_dot(_f(A), C.c_uint(A.shape[0]), C.c_uint(A.shape[1]), _f(B), C.c_uint(B.shape[0]),
                                C.c_uint(B.shape[1]), int(tA), int(tB), C.c_float(scale), _f(out))
        return out

    def adam(self, lr, dims):
        return _Adam(self.lib, "ref", lr, dims)

    # ---- the Lambda functions' own tensor ops (src/funcs/{gcn,gat}/ops), sequenced as their main.cpp does
    def funcs_gcn_forward(self, ah, W):
        """forwardLayer, funcs/gcn/main.cpp:244-250 -> (z, h)."""
        ah, W = _c32(ah), _c32(W)
        V, Fin = ah.shape
        z = np.empty((V, W.shape[1]), np.float32)
        h = np.empty_like(z)
        self.lib.ref_funcs_gcn_forward(_f(ah), _f(W), C.c_uint(V), C.c_uint(Fin), C.c_uint(W.shape[1]), _f(z), _f(h))
        return z, h

    def funcs_gcn_final(self, ah, W, lab, scale: float):
        """finalLayer, funcs/gcn/main.cpp:83-108.  `correct` / `loss` are checkAccuracy / checkLoss over
        ALL rows (forward_ops.cpp), `masked` the predictions after maskout()."""
        ah, W, lab = _c32(ah), _c32(W), _c32(lab)
        V, Fin = ah.shape
        Cc = W.shape[1]
        o = dict(pred=np.empty((V, Cc), np.float32), masked=np.empty((V, Cc), np.float32),
                 d=np.empty((V, Cc), np.float32), grad=np.empty((V, Fin), np.float32),
                 dW=np.empty((Fin, Cc), np.float32))
        correct, loss = C.c_uint(), C.c_float()
        self.lib.ref_funcs_gcn_final(_f(ah), _f(W), _f(lab), C.c_uint(V), C.c_uint(Fin), C.c_uint(Cc),
                                     C.c_float(scale), _f(o["pred"]), C.byref(correct), C.byref(loss),
                                     _f(o["masked"]), _f(o["d"]), _f(o["grad"]), _f(o["dW"]))
        o["correct"], o["loss"] = int(correct.value), float(loss.value)
        return o

    def funcs_gcn_backward(self, ah, z, aTg, W):
        """backwardLayer, funcs/gcn/main.cpp:176-189 -> (resultGrad, d_weights)."""
        ah, z, aTg, W = _c32(ah), _c32(z), _c32(aTg), _c32(W)
        V, Fin = ah.shape
        Fout = W.shape[1]
        grad = np.empty((V, Fin), np.float32)
        dW = np.empty((Fin, Fout), np.float32)
        self.lib.ref_funcs_gcn_backward(_f(ah), _f(z), _f(aTg), _f(W), C.c_uint(V), C.c_uint(Fin), C.c_uint(Fout),
                                        _f(grad), _f(dW))
        return grad, dW

    def funcs_gat_edge_forward(self, z, a, col_ptrs):
        """funcs/gat/main.cpp:84-94: az = edgeMatMul(eInfo, Z, a); A = leakyReLU(az)."""
        z, a = _c32(z), _c32(a).reshape(-1)
        V, F = z.shape
        ptrs = np.ascontiguousarray(col_ptrs, dtype=np.uint64)
        nnz = int(ptrs[V])
        az = np.zeros(max(nnz, 1), np.float32)
        A = np.zeros(max(nnz, 1), np.float32)
        self.lib.ref_funcs_gat_edge_forward(_f(z), _f(a), _q(ptrs), C.c_uint(V), C.c_uint(F), C.c_uint(nnz), _f(az), _f(A))
        return az[:nnz], A[:nnz]

    def funcs_gat_edge_backward(self, grad, az, z, a, col_ptrs):
        """funcs/gat/main.cpp:150-160: dA = expandHadamardMul(grad, leakyReLUDerivative(az)) . a -> (dA, dAct)."""
        grad, az, z, a = _c32(grad), _c32(az), _c32(z), _c32(a).reshape(-1)
        V, F = z.shape
        ptrs = np.ascontiguousarray(col_ptrs, dtype=np.uint64)
        nnz = int(ptrs[V])
        dA = np.zeros(max(nnz, 1), np.float32)
        dAct = np.zeros((max(nnz, 1), F), np.float32)
        self.lib.ref_funcs_gat_edge_backward(_f(grad), _f(az), _f(z), _f(a), _q(ptrs), C.c_uint(V), C.c_uint(F),
                                             C.c_uint(nnz), _f(dA), _f(dAct))
        return dA[:nnz], dAct[:nnz]

    def funcs_gat_expand_dot(self, m, v, col_ptrs):
        m, v = _c32(m), _c32(v).reshape(-1)
        V, F = m.shape
        ptrs = np.ascontiguousarray(col_ptrs, dtype=np.uint64)
        nnz = int(ptrs[V])
        out = np.zeros(max(nnz, 1), np.float32)
        self.lib.ref_funcs_gat_expand_dot(_f(m), _f(v), _q(ptrs), C.c_uint(V), C.c_uint(F), C.c_uint(nnz), _f(out))
        return out[:nnz]


class RefEngine:
    """The reference's OWN object code for the hot path (oracle/_ref/librefengine.so: gcn_ops.cpp and
    CPU_comm.cpp compiled in place, see oracle/ref_engine.cpp): Engine::preallocateGCN / aggregateGCN and
    CPUComm::NNCompute (vtxNNForwardGCN / vtxNNBackwardGCN) on one partition, driven chunk by chunk.
    Tensors are numpy views of the reference's own allocations (savedNNTensors[layer][name]).

    The weight server is an in-process stand-in: `set_weights` installs what getWeightMatrix hands out,
    `update(layer)` is what sendWeightUpdate received.  No exchange (single partition, or ghost blocks filled
    by the caller through `tensor(layer, "fg" | "bg")`)."""

    def __init__(self, graph_image, dims, gat: bool = False):
        import tempfile

        path = _build.build_ref_engine()
        if path is None or not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/librefengine.so unavailable (no /root/reference, no prebuilt)")
        self.lib = L = C.CDLL(path)
        L.refeng_create.restype = C.c_void_p
        L.refeng_tensor.restype = _f32p
        self.dims = [int(d) for d in dims]
        self.L = len(self.dims) - 1
        self._tmp = tempfile.TemporaryDirectory()
        gpath = os.path.join(self._tmp.name, "graph.0.bin")
        with open(gpath, "wb") as f:
            f.write(bytes(graph_image) if not isinstance(graph_image, np.ndarray) else graph_image.tobytes())
        wpath = os.path.join(self._tmp.name, "weightservers")
        with open(wpath, "w") as f:
            f.write("127.0.0.1\n")
        arr = (C.c_uint * len(self.dims))(*self.dims)
        self.gat = gat
        self.h = C.c_void_p(L.refeng_create(gpath.encode(), arr, C.c_uint(self.L), wpath.encode(), int(gat)))

    @staticmethod
    def available() -> bool:
        try:
            p = _build.build_ref_engine()
            return p is not None and os.path.exists(p)
        except Exception:
            return False

    def set_threads(self, n: int):
        self.lib.refeng_set_threads(int(n))

    def tensor(self, layer: int, name: str) -> np.ndarray:
        r, c = C.c_uint(), C.c_uint()
        p = self.lib.refeng_tensor(self.h, C.c_uint(layer), name.encode(), C.byref(r), C.byref(c))
        if not p:
            raise KeyError("savedNNTensors[%d][%r]" % (layer, name))
        return np.ctypeslib.as_array(p, shape=(int(r.value), int(c.value)))

    def set_weights(self, layer: int, w: np.ndarray):
        w = _c32(w)
        assert w.shape == (self.dims[layer], self.dims[layer + 1])
        self.lib.refeng_set_weights(self.h, C.c_uint(layer), _f(w))

    def update(self, layer: int) -> np.ndarray:
        dw = np.zeros((self.dims[layer], self.dims[layer + 1]), np.float32)
        if self.lib.refeng_get_update(self.h, C.c_uint(layer), _f(dw)) != 0:
            raise RuntimeError("no weight update was sent for layer %d yet" % layer)
        return dw

    def stats(self):
        a, l = C.c_float(), C.c_float()
        self.lib.refeng_stats(self.h, C.byref(a), C.byref(l))
        return float(a.value), float(l.value)

    def aggregate(self, layer: int, dir: int, low: int = 0, up=None):
        up = self.tensor(0, "x").shape[0] if up is None else up
        self.lib.refeng_aggregate(self.h, C.c_uint(layer), int(dir), C.c_uint(low), C.c_uint(up))

    def apply_vertex(self, layer: int, dir: int):
        self.lib.refeng_apply_vertex(self.h, C.c_uint(layer), int(dir))

    # ---- GAT (the reference's chunk.layer convention: feature layer + 1 for everything but ApplyVertex forward)
    def set_a(self, layer: int, a: np.ndarray):
        a = _c32(a).reshape(-1)
        assert a.size == self.dims[layer + 1]
        self.lib.refeng_set_a(self.h, C.c_uint(layer), _f(a))

    def a_update(self, layer: int) -> np.ndarray:
        da = np.zeros(self.dims[layer + 1], np.float32)
        if self.lib.refeng_get_a_update(self.h, C.c_uint(layer), _f(da)) != 0:
            raise RuntimeError("no a_i update was sent for layer %d yet" % layer)
        return da

    def gat_forward(self, l: int):
        """AV -> (SC) -> AE -> GA (-> predict on the last layer) for feature layer l."""
        self.lib.refeng_gat_op(self.h, 1, C.c_uint(l), 0)
        self.lib.refeng_gat_op(self.h, 2, C.c_uint(l + 1), 0)
        self.lib.refeng_gat_op(self.h, 0, C.c_uint(l + 1), 0)
        if l + 1 == self.L:
            self.lib.refeng_gat_op(self.h, 3, C.c_uint(l + 1), 0)

    def gat_backward(self, l: int):
        """(SC) -> AE -> GA -> AV backward for feature layer l."""
        self.lib.refeng_gat_op(self.h, 2, C.c_uint(l + 1), 1)
        self.lib.refeng_gat_op(self.h, 0, C.c_uint(l + 1), 1)
        self.lib.refeng_gat_op(self.h, 1, C.c_uint(l + 1), 1)

    def epoch_gcn(self):
        """One synchronous single-partition GCN epoch in the reference's operator order:
        GA -> AV forward per layer, then GA -> AV backward; weights are NOT stepped (the caller owns Adam)."""
        for l in range(self.L):
            self.aggregate(l, 0)
            self.apply_vertex(l, 0)
        for l in range(self.L - 1, 0, -1):
            self.aggregate(l, 1)
            self.apply_vertex(l, 1)
