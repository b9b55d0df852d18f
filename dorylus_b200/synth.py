"""Synthetic datasets in the shapes BASELINE.json names (no network: no real Reddit / Cora).

Generators are ours, formats are the reference's (dorylus_b200/formats.py).  Seeds are fixed so a
config is reproducible from its name.  Feature / label recipes follow the reference's own
generators: inputs/generateFeatues.cpp:32-55 (per row between F/3 and 3F/4 non-zeros ~U(-1,1)) and
inputs/generateLabels.cpp (uniform class ids; we draw from [0, kinds) so readLabelsFile's
``label < labelKinds`` assert, engine/utils.cpp:582, holds).
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class GraphSpec:
    name: str
    num_vertices: int
    num_edges: int  # directed in-edge records (both directions present), self loops excluded
    dims: list
    seed: int
    sigma: float = 1.0  # log-normal spread of the expected degrees
    locality: float = 0.0  # probability that an edge stays inside its source's community block
    communities: int = 1


CONFIGS = {
    # BASELINE.json configs[0]: Cora GCN 2-layer (2.7K verts, 10K edges, 1433 feat)
    "cora": GraphSpec("cora", 2708, 10556, [1433, 16, 7], seed=1, sigma=0.6),
    # BASELINE.json configs[1]/[2]: Reddit (232K verts, 114M edges, 602 feat), dims run/reddit.config
    "reddit": GraphSpec("reddit", 232965, 114615892, [602, 128, 41], seed=11, sigma=1.0),
    # the same degree sequence with community structure (not a BASELINE config; DESIGN.md section 11:
    # what a tile-reuse aggregation kernel would be measured on): 80 % of the edges stay inside one
    # of 1,024 contiguous communities of ~227 vertices
    "reddit-communities": GraphSpec("reddit-communities", 232965, 114615892, [602, 128, 41], seed=11, sigma=1.0,
                                    locality=0.8, communities=1024),
    # scaled-down Reddit shape for parity tests that run the CPU oracle in seconds
    "reddit-small": GraphSpec("reddit-small", 8192, 8192 * 96, [602, 128, 41], seed=21, sigma=1.0),
    "reddit-tiny": GraphSpec("reddit-tiny", 600, 600 * 24, [602, 128, 41], seed=31, sigma=0.8),
    # BASELINE.json configs[3]: Amazon GCN 3-layer, synthetic 100-dim features (V, E are our choice)
    "amazon": GraphSpec("amazon", 9_430_088, 231_594_310, [100, 64, 64, 25], seed=41, sigma=0.9,
                        locality=0.9, communities=4096),
    "amazon-tiny": GraphSpec("amazon-tiny", 900, 900 * 20, [100, 64, 64, 25], seed=43, sigma=0.9, locality=0.9, communities=16),
    # BASELINE.json configs[4]: Friendster GCN 2-layer (65M verts, 1.8B edges, 16-dim features)
    "friendster": GraphSpec("friendster", 65_608_366, 1_800_000_000, [16, 48, 51], seed=51, sigma=0.9,
                            locality=0.9, communities=32768),
}


def _alias_table(w: np.ndarray):
    """Walker alias table for O(1) sampling from weights w (vectorised two-stack build)."""
    n = w.size
    p = w.astype(np.float64) * (n / w.sum())
    alias = np.arange(n, dtype=np.int64)
    small = list(np.nonzero(p < 1.0)[0])
    large = list(np.nonzero(p >= 1.0)[0])
    p = p.copy()
    while small and large:
        s = small.pop()
        l = large[-1]
        alias[s] = l
        p[l] -= 1.0 - p[s]
        if p[l] < 1.0:
            large.pop()
            small.append(l)
    prob = np.minimum(p, 1.0)
    for i in small + large:
        prob[i] = 1.0
    return prob.astype(np.float32), alias.astype(np.uint32)


def _alias_sample(rng, prob, alias, n):
    idx = rng.integers(0, prob.size, size=n, dtype=np.uint32)
    keep = rng.random(n, dtype=np.float32) < prob[idx]
    return np.where(keep, idx, alias[idx]).astype(np.uint32)


def generate_edges(spec: GraphSpec):
    """Undirected Chung-Lu graph with log-normal expected degrees, written as two directed records
    per undirected edge (the shipped datasets carry both directions, run/run-onnode:46).
    Self loops are not generated; duplicate edges are kept (the reference keeps them, Q2).
    With `locality` > 0 an edge's second endpoint is drawn from the same contiguous community
    block as the first with that probability (H5: graphs without locality make every remote
    vertex a ghost)."""
    rng = np.random.default_rng(spec.seed)
    V = spec.num_vertices
    n_und = spec.num_edges // 2
    if V > 2_000_000 and spec.locality == 0:
        return _configuration_model(rng, V, n_und, spec.sigma)
    w = np.exp(spec.sigma * rng.standard_normal(V)).astype(np.float64)
    prob, alias = _alias_table(w)
    u = _alias_sample(rng, prob, alias, n_und)
    v = _alias_sample(rng, prob, alias, n_und)
    if spec.locality > 0 and spec.communities > 1:
        blk = (V + spec.communities - 1) // spec.communities
        local = rng.random(n_und, dtype=np.float32) < spec.locality
        base = (u // blk) * blk
        span = np.minimum(base + blk, V) - base
        v_loc = base + (v % span)
        v = np.where(local, v_loc, v).astype(np.uint32)
    # remove self loops by nudging the second endpoint (keeps the edge count exact)
    same = u == v
    v = np.where(same, (v + 1) % V, v).astype(np.uint32)
    src = np.empty(2 * n_und, dtype=np.uint32)
    dst = np.empty(2 * n_und, dtype=np.uint32)
    src[0::2], dst[0::2] = u, v
    src[1::2], dst[1::2] = v, u
    return src, dst


def _configuration_model(rng, V: int, n_und: int, sigma: float):
    """Large graphs: log-normal degree sequence + random stub matching (O(E), no per-vertex Python
    loop).  Same statistics as the Chung-Lu sampler above up to the exactness of the degrees."""
    w = np.exp(sigma * rng.standard_normal(V))
    deg = np.floor(w * (2.0 * n_und / w.sum())).astype(np.int64)
    short = 2 * n_und - int(deg.sum())
    bump = rng.integers(0, V, size=short)
    np.add.at(deg, bump, 1)
    stubs = np.repeat(np.arange(V, dtype=np.uint32), deg)
    rng.shuffle(stubs)
    u, v = stubs[:n_und], stubs[n_und:2 * n_und]
    v = np.where(u == v, (v + 1) % V, v).astype(np.uint32)
    src = np.empty(2 * n_und, dtype=np.uint32)
    dst = np.empty(2 * n_und, dtype=np.uint32)
    src[0::2], dst[0::2] = u, v
    src[1::2], dst[1::2] = v, u
    return src, dst


def generate_features(num_vertices: int, dim: int, seed: int, dense: bool = True) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.random((num_vertices, dim), dtype=np.float32) * 2.0 - 1.0
    if not dense:  # inputs/generateFeatues.cpp:32-55
        k = rng.integers(dim // 3, max(dim * 3 // 4, dim // 3 + 1), size=num_vertices)
        r = rng.random((num_vertices, dim), dtype=np.float32)
        thresh = (k / float(dim)).astype(np.float32)[:, None]
        x = np.where(r < thresh, x, 0.0).astype(np.float32)
    return x


def generate_labels(num_vertices: int, kinds: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, kinds, size=num_vertices, dtype=np.uint32)


def contiguous_parts(num_vertices: int, num_parts: int) -> np.ndarray:
    """Equal contiguous vertex ranges (what a locality-preserving partitioner yields on a
    community-ordered graph; METIS, inputs/partitioner.cpp:113, is not available here)."""
    blk = (num_vertices + num_parts - 1) // num_parts
    return (np.arange(num_vertices, dtype=np.int64) // blk).astype(np.int32)


def random_parts(num_vertices: int, num_parts: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, num_parts, size=num_vertices).astype(np.int32)


def edge_cut(src, dst, parts) -> float:
    return float(np.mean(parts[src] != parts[dst]))
