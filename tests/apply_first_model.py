"""CPU model of the engine's apply-first schedule (DORY_FLAG_APPLY_FIRST), test infrastructure.

The reference computes every GCN layer as aggregate -> apply: ah = A_hat . in, z = ah . W
(engine/ops/gcn_ops.cpp:130-191, commmanager/CPU_comm.cpp:98-159).  Where a layer narrows
(F_out < F_in) the engine can run it apply -> aggregate instead -- A_hat . (in . W) -- which gathers
F_out-wide rows instead of F_in-wide ones.  This module states that schedule on the oracle's own
operators (aggregation = oracle.aggregate_gcn, dense products in float32) for ALL partitions of a
graph, so that tests can check (1) on the CPU that it reproduces the reference epoch's tensors, and
(2) on the GPU that the engine's apply-first path reproduces the REFERENCE oracle.

Per apply-first layer l (in = x or h[l-1], local rows only):
    forward   t = in . W ; exchange t's ghost rows (forward send lists) ; z = A_hat [t; t_ghost]
              h = tanh(z)                       (last layer: soft-max / maskout / scale on z -> g)
    backward  g = dL/dz (hidden: aTg[l] (*) (1 - h^2)) ; exchange g's ghost rows (backward lists)
              u = A_hat^T [g; g_ghost] ; dW = in^T . u ; aTg[l-1] = u . W^T   (= dL/dh[l-1], l > 0)
A standard layer keeps the reference's order.  aTg[l-1] always means dL/dh[l-1].
"""
from __future__ import annotations

import numpy as np

from oracle.driver import OracleGCN


def choose_apply_first(dims, padded=lambda f: (f + 31) // 32 * 32 if f > 16 else (f + 3) // 4 * 4):
    """The engine's rule: a layer runs apply-first when its output rows are narrower in HBM than its
    input rows (padded pitch, dorylus_b200/csrc/common.cuh: padded_ld)."""
    return [padded(dims[l + 1]) < padded(dims[l]) for l in range(len(dims) - 1)]


class ApplyFirstGCN(OracleGCN):
    def __init__(self, oracle, graphs, dims, lr: float = 0.01, apply_first=None):
        super().__init__(oracle, graphs, dims, lr)
        self.af = list(apply_first) if apply_first is not None else choose_apply_first(dims)
        for p, g in enumerate(graphs):
            V, Gs, Gd = g.local_vtx_cnt, g.src_ghost_cnt, g.dst_ghost_cnt
            for l in range(self.L):
                if self.af[l]:
                    t = self.saved[p][l]
                    F = dims[l + 1]
                    t["t"], t["fg_t"] = np.zeros((V, F), np.float32), np.zeros((Gs, F), np.float32)
                    t["g"], t["bg_g"] = np.zeros((V, F), np.float32), np.zeros((Gd, F), np.float32)
                    t["u"] = np.zeros((V, F), np.float32)
                    t.setdefault("z", np.zeros((V, F), np.float32))

    def _exchange(self, layer, local_name, ghost_name, forward: bool):
        for p, g in enumerate(self.graphs):
            lists = g.fwd_send if forward else g.bwd_send
            for q in range(self.P):
                if q == p or lists[q].size == 0:
                    continue
                gq = self.graphs[q]
                gv = g.local_to_global[lists[q]]
                ghosts = gq.src_ghost_gvid if forward else gq.dst_ghost_gvid
                slots = np.searchsorted(ghosts, gv)
                assert np.array_equal(ghosts[slots], gv)
                self.saved[q][layer][ghost_name][slots] = self.saved[p][layer][local_name][lists[q]]

    def _input(self, p, l):
        return self.saved[p][0]["x"] if l == 0 else self.saved[p][l - 1]["h"]

    def _last_layer_gradient(self, p, z):
        """softmax -> getTrainStat -> maskout (floats, Q6) -> (P - Y) / (V_global * 0.66), as
        CPU_comm.cpp:108-121; computed through the oracle by handing it z as `ah` with W = I."""
        g, t = self.graphs[p], self.saved[p][self.L - 1]
        C = z.shape[1]
        r = self.o.vtx_forward_gcn_last(z, np.eye(C, dtype=np.float32), t["lab"], g.global_vtx_cnt)
        self.acc[p], self.loss[p] = r["acc"], r["loss"]
        return r["d"]

    def epoch(self):
        L, o = self.L, self.o
        for l in range(L):
            last = l == L - 1
            if self.af[l]:
                for p in range(self.P):
                    self.saved[p][l]["t"][:] = o.dot(self._input(p, l), self.W[l])
                self._exchange(l, "t", "fg_t", True)
                for p, g in enumerate(self.graphs):
                    t = self.saved[p][l]
                    o.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, t["t"], t["fg_t"], out=t["z"])
                    if last:
                        t["g"][:] = self._last_layer_gradient(p, t["z"])
                    else:
                        t["h"][:] = np.tanh(t["z"])
            else:
                for p in range(self.P):
                    self.aggregate(p, l, 0)
                    self.apply_vertex_forward(p, l)  # last layer: grad[l] = d . W^T and dW[l]
            if not last and not self.af[l + 1]:
                self.scatter(l + 1, 0)
        for l in range(L - 1, -1, -1):
            if self.af[l]:
                self._exchange(l, "g", "bg_g", False)
                for p, g in enumerate(self.graphs):
                    t = self.saved[p][l]
                    o.aggregate_gcn(g.row_ptrs, g.col_idxs, g.bwd_vals, g.norms, t["g"], t["bg_g"], out=t["u"])
                    self.dW[p][l] = o.dot(self._input(p, l), t["u"], tA=True)
                    if l > 0:
                        self.saved[p][l - 1]["aTg"][:] = o.dot(t["u"], self.W[l], tB=True)
            elif l > 0:
                self.scatter(l, 1)
                for p in range(self.P):
                    self.aggregate(p, l, 1)
            if l == 0:
                break
            for p in range(self.P):  # dL/dz of layer l-1 from dL/dh[l-1] = aTg[l-1]
                t = self.saved[p][l - 1]
                if self.af[l - 1]:
                    t["g"][:] = t["aTg"] * (1 - t["h"].astype(np.float32) ** 2)
                else:
                    self.apply_vertex_backward(p, l - 1)
        self.apply_updates()
        return dict(acc=list(self.acc), loss=list(self.loss))
