"""Synchronous driver of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Runs the reference's per-chunk state machine (SURVEY.md §3.1 for GCN, §3.4 for GAT) for ALL
partitions of a graph inside one process: Gather / ApplyVertex / Scatter / ApplyEdge are the
oracle functions of oracle/oracle.cpp, the ghost exchange is the receiver loop of
ghostReceiverGCN (engine/ops/gcn_ops.cpp:310-318: rows addressed by global id through the ghost
map), the weight server is xavier + sync Adam with the gradient summed over partitions
(weight-server/weighttensor.cpp:263-284).  No ZeroMQ, no threads, no Lambda.

Tensor names and layer indices are exactly those of Engine::savedNNTensors.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from dorylus_b200.formats import PartitionGraph

from .pyoracle import Oracle

FORWARD, BACKWARD = 0, 1
TRAIN_PORTION, VAL_PORTION = 0.66, 0.1


class OracleGCN:
    def __init__(self, oracle: Oracle, graphs: List[PartitionGraph], dims: List[int], lr: float = 0.01):
        self.o, self.graphs, self.dims, self.L = oracle, graphs, list(dims), len(dims) - 1
        self.P = len(graphs)
        self.saved: List[Dict[int, Dict[str, np.ndarray]]] = []
        self._tables = {}
        for g in graphs:
            V, Gs, Gd = g.local_vtx_cnt, g.src_ghost_cnt, g.dst_ghost_cnt
            t: Dict[int, Dict[str, np.ndarray]] = {l: {} for l in range(self.L)}
            t[0]["x"] = np.zeros((V, dims[0]), np.float32)
            t[0]["fg"] = np.zeros((Gs, dims[0]), np.float32)
            t[self.L - 1]["lab"] = np.zeros((V, dims[self.L]), np.float32)
            for l in range(self.L):  # preallocateGCN, gcn_ops.cpp:43-69
                t[l]["ah"] = np.zeros((V, dims[l]), np.float32)
                if l < self.L - 1:
                    t[l]["z"] = np.zeros((V, dims[l + 1]), np.float32)
                    t[l]["h"] = np.zeros((V, dims[l + 1]), np.float32)
                    t[l + 1]["fg"] = np.zeros((Gs, dims[l + 1]), np.float32)
            for l in range(self.L - 1, 0, -1):  # gcn_ops.cpp:72-92
                t[l]["grad"] = np.zeros((V, dims[l]), np.float32)
                t[l - 1]["bg"] = np.zeros((Gd, dims[l]), np.float32)
                t[l - 1]["aTg"] = np.zeros((V, dims[l]), np.float32)
            self.saved.append(t)
        self.W = [oracle.xavier(dims[l], dims[l + 1]) for l in range(self.L)]
        self.dW = [[None] * self.L for _ in graphs]
        self.adam = oracle.adam(lr, dims)
        self.acc = [0.0] * self.P
        self.loss = [0.0] * self.P

    # ------------------------------------------------------------------ data
    def load_features(self, feats: np.ndarray, labels_onehot: np.ndarray):
        """readFeaturesFile / readLabelsFile (engine/utils.cpp:486-596)."""
        for p, g in enumerate(self.graphs):
            self.saved[p][0]["x"][:] = feats[g.local_to_global]
            self.saved[p][0]["fg"][:] = feats[g.src_ghost_gvid]
            self.saved[p][self.L - 1]["lab"][:] = labels_onehot[g.local_to_global]

    # ------------------------------------------------------------------ operators
    def _table(self, p: int, layer: int, dir: int, ptrs, idxs, local, ghost):
        """savedEdgeTensors[layer]["fedge" | "bedge"]: the reference builds every pointer table ONCE, in
        preallocateGCN (gcn_ops.cpp:38,66,86); the tensors they point into never move (updated in place)."""
        key = (p, layer, dir)
        if key not in self._tables:
            self._tables[key] = self.o.edge_table(ptrs, idxs, local, ghost)
        return self._tables[key]

    def aggregate(self, p: int, layer: int, dir: int, low: int = 0, up=None):
        g, t = self.graphs[p], self.saved[p]
        if dir == FORWARD:  # gcn_ops.cpp:139-147
            local = t[0]["x"] if layer == 0 else t[layer - 1]["h"]
            tab = self._table(p, layer, dir, g.col_ptrs, g.row_idxs, local, t[layer]["fg"])
            self.o.aggregate_gcn(g.col_ptrs, g.row_idxs, g.fwd_vals, g.norms, local, t[layer]["fg"],
                                 low, up, out=t[layer]["ah"], table=tab)
        else:  # :148-154
            tab = self._table(p, layer, dir, g.row_ptrs, g.col_idxs, t[layer]["grad"], t[layer - 1]["bg"])
            self.o.aggregate_gcn(g.row_ptrs, g.col_idxs, g.bwd_vals, g.norms, t[layer]["grad"], t[layer - 1]["bg"],
                                 low, up, out=t[layer - 1]["aTg"], table=tab)

    def apply_vertex_forward(self, p: int, layer: int):
        g, t = self.graphs[p], self.saved[p]
        if layer < self.L - 1:
            z, h = self.o.vtx_forward_gcn_hidden(t[layer]["ah"], self.W[layer])
            t[layer]["z"][:], t[layer]["h"][:] = z, h
        else:
            r = self.o.vtx_forward_gcn_last(t[layer]["ah"], self.W[layer], t[layer]["lab"], g.global_vtx_cnt)
            t[layer]["grad"][:] = r["grad"]
            self.dW[p][layer] = r["dW"]
            self.acc[p], self.loss[p] = r["acc"], r["loss"]
            self.last = r

    def apply_vertex_backward(self, p: int, layer: int):
        t = self.saved[p]
        dW, grad = self.o.vtx_backward_gcn(t[layer]["aTg"], t[layer]["z"], t[layer]["ah"], self.W[layer], layer != 0)
        self.dW[p][layer] = dW
        if layer != 0:
            t[layer]["grad"][:] = grad

    def scatter(self, layer: int, dir: int):
        """scatterGCN + ghostReceiverGCN for every (sender, receiver) pair."""
        for p, g in enumerate(self.graphs):
            src = self.saved[p][layer - 1]["h"] if dir == FORWARD else self.saved[p][layer]["grad"]
            lists = g.fwd_send if dir == FORWARD else g.bwd_send
            for q in range(self.P):
                if q == p or lists[q].size == 0:
                    continue
                gq = self.graphs[q]
                gv = g.local_to_global[lists[q]]
                if dir == FORWARD:
                    slots = np.searchsorted(gq.src_ghost_gvid, gv)
                    assert np.array_equal(gq.src_ghost_gvid[slots], gv)
                    self.saved[q][layer]["fg"][slots] = src[lists[q]]
                else:
                    slots = np.searchsorted(gq.dst_ghost_gvid, gv)
                    assert np.array_equal(gq.dst_ghost_gvid[slots], gv)
                    self.saved[q][layer - 1]["bg"][slots] = src[lists[q]]

    def apply_updates(self):
        for l in range(self.L - 1, -1, -1):
            total = self.dW[0][l].copy()
            for p in range(1, self.P):
                total += self.dW[p][l]  # lPtr[u] += gPtr[u], weighttensor.cpp:266-267
            self.adam.update(l, self.W[l], total)

    # ------------------------------------------------------------------ epoch
    def epoch(self):
        L = self.L
        for l in range(L):
            for p in range(self.P):
                self.aggregate(p, l, FORWARD)
                self.apply_vertex_forward(p, l)
            if l < L - 1:
                self.scatter(l + 1, FORWARD)
        self.scatter(L - 1, BACKWARD)
        for l in range(L - 1, 0, -1):
            for p in range(self.P):
                self.aggregate(p, l, BACKWARD)
                self.apply_vertex_backward(p, l - 1)
            if l - 1 > 0:
                self.scatter(l - 1, BACKWARD)
        self.apply_updates()
        return dict(acc=list(self.acc), loss=list(self.loss))


class OracleGAT:
    """The GAT chunk state machine of SURVEY.md §3.4 (fixed weights: quirk Q10)."""

    def __init__(self, oracle: Oracle, graphs: List[PartitionGraph], dims: List[int], predict_from: str = "az"):
        self.o, self.graphs, self.dims, self.L = oracle, graphs, list(dims), len(dims) - 1
        self.P = len(graphs)
        self.predict_from = predict_from
        self.saved = []
        self.A = []  # forwardAdj.values alias, shared by all layers (Q12)
        for g in graphs:
            V, Gs, Gd, E = g.local_vtx_cnt, g.src_ghost_cnt, g.dst_ghost_cnt, g.local_in_edge_cnt
            t = {l: {} for l in range(self.L)}
            t[0]["h"] = np.zeros((V, dims[0]), np.float32)
            t[self.L - 1]["lab"] = np.zeros((V, dims[self.L]), np.float32)
            for l in range(self.L):
                nf = dims[l + 1]
                t[l]["z"] = np.zeros((V, nf), np.float32)
                t[l]["fg_z"] = np.zeros((Gs, nf), np.float32)
                t[l]["az"] = np.zeros(E, np.float32)
                t[l]["ah"] = np.zeros((V, nf), np.float32)
                t[l]["grad"] = np.zeros((V, nf), np.float32)
                t[l]["bg_d"] = np.zeros((Gd, nf), np.float32)
                t[l]["dA"] = np.zeros(E, np.float32)
                t[l]["aTg"] = np.zeros((V, nf), np.float32)
            self.saved.append(t)
            self.A.append(np.zeros(E, np.float32))
        self.W = [oracle.xavier(dims[l], dims[l + 1]) for l in range(self.L)]
        self.a = [oracle.kaiming(dims[l + 1], 1) for l in range(self.L)]
        self.dW = [[None] * self.L for _ in graphs]
        self.da = [[None] * self.L for _ in graphs]

    def load_features(self, feats, labels_onehot):
        for p, g in enumerate(self.graphs):
            self.saved[p][0]["h"][:] = feats[g.local_to_global]
            self.saved[p][self.L - 1]["lab"][:] = labels_onehot[g.local_to_global]

    def _scatter(self, fl: int, dir: int):
        for p, g in enumerate(self.graphs):
            src = self.saved[p][fl]["z"] if dir == FORWARD else self.saved[p][fl]["grad"]
            lists = g.fwd_send if dir == FORWARD else g.bwd_send
            for q in range(self.P):
                if q == p or lists[q].size == 0:
                    continue
                gq = self.graphs[q]
                gv = g.local_to_global[lists[q]]
                if dir == FORWARD:
                    self.saved[q][fl]["fg_z"][np.searchsorted(gq.src_ghost_gvid, gv)] = src[lists[q]]
                else:
                    self.saved[q][fl]["bg_d"][np.searchsorted(gq.dst_ghost_gvid, gv)] = src[lists[q]]

    def forward_layer(self, l: int):
        for p, g in enumerate(self.graphs):
            t = self.saved[p]
            feats = t[0]["h"] if l == 0 else t[l - 1]["ah"]
            t[l]["z"][:] = self.o.vtx_forward_gat(feats, self.W[l])
        self._scatter(l, FORWARD)
        for p, g in enumerate(self.graphs):
            t = self.saved[p]
            az, A = self.o.edg_forward_gat(t[l]["z"], self.a[l], g.col_ptrs)
            t[l]["az"][:], self.A[p][:] = az, A
            self.o.aggregate_gat_fwd(g.col_ptrs, g.row_idxs, self.A[p], t[l]["z"], t[l]["fg_z"], out=t[l]["ah"])
            if l == self.L - 1:
                C = self.dims[self.L]
                V = g.local_vtx_cnt
                if self.predict_from == "az":
                    logits = t[l]["az"][: V * C].reshape(V, C)  # quirk Q9
                else:
                    logits = t[l]["ah"]
                t[l]["grad"][:] = self.o.predict_gat(logits, t[l]["lab"])

    def backward_layer(self, l: int):
        self._scatter(l, BACKWARD)
        for p, g in enumerate(self.graphs):
            t = self.saved[p]
            dA, da = self.o.edg_backward_gat(t[l]["grad"], t[l]["az"], t[l]["z"], self.a[l], g.col_ptrs)
            t[l]["dA"][:] = dA
            self.da[p][l] = da
            self.o.aggregate_gat_bwd(g.col_ptrs, g.row_idxs, t[l]["dA"], t[l]["z"], t[l]["fg_z"], g.row_ptrs,
                                     g.col_idxs, g.bwd_vals, t[l]["grad"], t[l]["bg_d"], out=t[l]["aTg"])
            feats = t[0]["h"] if l == 0 else t[l - 1]["ah"]
            dW, grad = self.o.vtx_backward_gat(feats, t[l]["aTg"], self.W[l], l != 0)
            self.dW[p][l] = dW
            if l != 0:
                t[l - 1]["grad"][:] = grad

    def epoch(self):
        for l in range(self.L):
            self.forward_layer(l)
        for l in range(self.L - 1, -1, -1):
            self.backward_layer(l)
