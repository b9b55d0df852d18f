"""Shared builders for the parity tests (seeded synthetic datasets, partition graphs, metrics)."""
from __future__ import annotations

import numpy as np

from dorylus_b200 import engine as dengine
from dorylus_b200 import formats, synth


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """Norm-wise relative error max|a-b| / max|b| (SURVEY.md §8d parity metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b)) if b.size else 0.0
    if denom == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / denom)


class Dataset:
    def __init__(self, V, src, dst, dims, feats, labels, parts, P, undirected=False):
        self.V, self.src, self.dst, self.dims = V, src, dst, list(dims)
        self.feats, self.labels, self.parts, self.P = feats, labels, parts, P
        self.onehot = formats.one_hot(labels, dims[-1])
        self.images = [dengine.preprocess_edges(src, dst, parts, V, p, P, undirected) for p in range(P)]
        self.graphs = [formats.parse_graph_bin(im) for im in self.images]


def random_dataset(V, E_und, dims, P=1, seed=0, parts="random", sigma=0.8, dense_feats=True, extra_edges=None):
    spec = synth.GraphSpec("t", V, 2 * E_und, list(dims), seed=seed, sigma=sigma)
    src, dst = synth.generate_edges(spec)
    if extra_edges is not None:
        src = np.concatenate([src, extra_edges[0].astype(np.uint32)])
        dst = np.concatenate([dst, extra_edges[1].astype(np.uint32)])
    feats = synth.generate_features(V, dims[0], seed + 1, dense=dense_feats)
    labels = synth.generate_labels(V, dims[-1], seed + 2)
    if P == 1:
        pr = np.zeros(V, np.int32)
    elif parts == "random":
        pr = synth.random_parts(V, P, seed + 3)
    else:
        pr = synth.contiguous_parts(V, P)
    return Dataset(V, src, dst, dims, feats, labels, pr, P)


def dense_normalized_adjacency(V, src, dst):
    """A_hat = D^-1/2 (A) D^-1/2 + D^-1 with A[dst, src] += 1 per record (duplicates kept, self loops
    dropped), D = in-degree + 1 -- the dense statement the aggregation must equal (numpy-gnn
    layers.py:199-210 restricted to symmetric inputs)."""
    keep = src != dst
    s, d = src[keep].astype(np.int64), dst[keep].astype(np.int64)
    A = np.zeros((V, V), np.float64)
    np.add.at(A, (d, s), 1.0)
    deg = np.bincount(d, minlength=V).astype(np.float64) + 1.0
    dinv = deg ** -0.5
    return A * dinv[:, None] * dinv[None, :] + np.diag(1.0 / deg)
