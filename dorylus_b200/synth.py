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


# ---------------------------------------------------------------------------- per-rank (streamed) generation
# The Amazon / Friendster shapes (BASELINE.json configs[3], [4]) are run on 8 GPUs with one process per
# GPU; a 1.8 G-record edge list (14 GB) must never exist in one process.  The generator below is a
# pure function of (spec, chunk): every rank walks the same chunk sequence and keeps the records
# incident to its own vertex range, so the union over ranks is one consistent graph and a rank's list
# equals the whole list filtered by incidence, in the same order (tests/test_synth_streamed.py).
#
# Degrees: inside a community block of B vertices, position i has weight
#   w(i) = mass of N(sigma, 1) on the i-th 1/B-quantile cell of N(0, 1)
# i.e. the size-biased (edge-endpoint) distribution of log-normal(sigma) weights assigned by rank, so an
# endpoint is drawn without any table: z ~ N(sigma, 1), i = floor(Phi(z) * B).  A fixed affine map
# scatters the ranks over the block's ids.  Every block receives the same number of first endpoints;
# the second endpoint stays in the block with probability `locality`, else it falls into a uniformly
# drawn block.
_CHUNK_EDGES = 2_000_000  # undirected edges per generation chunk (target)
_PERM_MUL = 1_000_003     # affine scatter of degree ranks inside a block (prime; coprime to any block size here)


class StreamedLayout:
    """Chunking of a community-structured GraphSpec: blocks of `blk` vertices, chunks of `bpc` blocks."""

    def __init__(self, spec: GraphSpec):
        V = spec.num_vertices
        self.V, self.n_und = V, spec.num_edges // 2
        self.blk = (V + spec.communities - 1) // max(spec.communities, 1)
        self.nblocks = (V + self.blk - 1) // self.blk
        target = max(1, self.n_und // _CHUNK_EDGES)
        self.bpc = (self.nblocks + target - 1) // target  # blocks per chunk
        self.nchunks = (self.nblocks + self.bpc - 1) // self.bpc
        # undirected edges whose first endpoint lies in chunk c: proportional to its block count
        cum = (np.minimum(np.arange(self.nchunks + 1, dtype=np.int64) * self.bpc, self.nblocks) * self.n_und) // self.nblocks
        self.quota = np.diff(cum)
        self.n_local = (self.quota * spec.locality).astype(np.int64)  # stay inside the block
        self.n_remote = self.quota - self.n_local

    def chunk_vertex_range(self, c: int):
        return c * self.bpc * self.blk, min((c + 1) * self.bpc * self.blk, self.V)


def _draw_in_blocks(rng, blocks: np.ndarray, lay: StreamedLayout, sigma: float) -> np.ndarray:
    from scipy.special import ndtr

    base = blocks * lay.blk
    size = np.minimum(base + lay.blk, lay.V) - base
    q = ndtr(rng.standard_normal(blocks.size) + sigma)
    i = np.minimum((q * size).astype(np.int64), size - 1)
    return base + (i * _PERM_MUL + 7) % size


def _chunk_edges(spec: GraphSpec, lay: StreamedLayout, c: int, part: int):
    """Undirected edges (u, v) of chunk c: part 0 = both endpoints in one block, part 1 = second endpoint
    in a uniformly drawn block.  Deterministic in (spec.seed, c, part)."""
    n = int(lay.n_local[c] if part == 0 else lay.n_remote[c])
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    rng = np.random.default_rng([spec.seed, c, part])
    b0, b1 = c * lay.bpc, min((c + 1) * lay.bpc, lay.nblocks)
    ub = rng.integers(b0, b1, size=n, dtype=np.int64)
    u = _draw_in_blocks(rng, ub, lay, spec.sigma)
    vb = ub if part == 0 else rng.integers(0, lay.nblocks, size=n, dtype=np.int64)
    v = _draw_in_blocks(rng, vb, lay, spec.sigma)
    # no self loops: move the second endpoint to the next vertex of its block (a block has >= 2 vertices)
    base = vb * lay.blk
    size = np.minimum(base + lay.blk, lay.V) - base
    v = np.where(u == v, base + (v - base + 1) % size, v)
    return u, v


def generate_incident_edges(spec: GraphSpec, rank: int, world: int, threads: int = 1):
    """Edge records (both directions of every undirected edge) incident to the vertex range of `rank`
    under contiguous_parts(V, world), in the order of the whole graph's record list; world = 1 gives the
    whole list.  Returns (src, dst, in_degree_of_own_vertices, (lo, hi))."""
    import concurrent.futures as cf

    lay = StreamedLayout(spec)
    V = lay.V
    pblk = (V + world - 1) // world
    lo, hi = min(rank * pblk, V), min((rank + 1) * pblk, V)

    def work(c: int):
        out = []
        c_lo, c_hi = lay.chunk_vertex_range(c)
        for part in (0, 1):
            if part == 0 and (c_hi <= lo or c_lo >= hi):
                continue  # both endpoints inside the chunk: nothing incident to this rank
            u, v = _chunk_edges(spec, lay, c, part)
            if world > 1:
                keep = ((u >= lo) & (u < hi)) | ((v >= lo) & (v < hi))
                u, v = u[keep], v[keep]
            s = np.empty(2 * u.size, np.uint32)
            d = np.empty(2 * u.size, np.uint32)
            s[0::2], d[0::2] = u, v
            s[1::2], d[1::2] = v, u
            out.append((s, d))
        return out

    if threads > 1:
        with cf.ThreadPoolExecutor(max_workers=threads) as ex:
            pieces = list(ex.map(work, range(lay.nchunks)))
    else:
        pieces = [work(c) for c in range(lay.nchunks)]
    flat = [p for ps in pieces for p in ps]
    src = np.concatenate([p[0] for p in flat]) if flat else np.zeros(0, np.uint32)
    dst = np.concatenate([p[1] for p in flat]) if flat else np.zeros(0, np.uint32)
    own = (dst >= lo) & (dst < hi)
    deg = np.bincount(dst[own].astype(np.int64) - lo, minlength=hi - lo).astype(np.uint32)
    return src, dst, deg, (lo, hi)


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


def generate_feature_rows(lo: int, hi: int, dim: int, seed: int, rows_per_chunk: int = 1 << 16) -> np.ndarray:
    """Dense U(-1, 1) feature rows of the global vertices [lo, hi), identical whatever range is asked
    for: rows are drawn in fixed chunks of `rows_per_chunk` vertices, each from its own stream."""
    out = np.empty((hi - lo, dim), np.float32)
    for c in range(lo // rows_per_chunk, (max(hi, lo + 1) - 1) // rows_per_chunk + 1):
        a, b = c * rows_per_chunk, (c + 1) * rows_per_chunk
        rows = np.random.default_rng([seed, c]).random((rows_per_chunk, dim), dtype=np.float32) * 2.0 - 1.0
        s, e = max(a, lo), min(b, hi)
        if s < e:
            out[s - lo:e - lo] = rows[s - a:e - a]
    return out


def generate_label_rows(lo: int, hi: int, kinds: int, seed: int, rows_per_chunk: int = 1 << 20) -> np.ndarray:
    out = np.empty(hi - lo, np.uint32)
    for c in range(lo // rows_per_chunk, (max(hi, lo + 1) - 1) // rows_per_chunk + 1):
        a, b = c * rows_per_chunk, (c + 1) * rows_per_chunk
        lab = np.random.default_rng([seed, c]).integers(0, kinds, size=rows_per_chunk, dtype=np.uint32)
        s, e = max(a, lo), min(b, hi)
        if s < e:
            out[s - lo:e - lo] = lab[s - a:e - a]
    return out


def contiguous_parts(num_vertices: int, num_parts: int) -> np.ndarray:
    """Equal contiguous vertex ranges (what a locality-preserving partitioner yields on a
    community-ordered graph; METIS, inputs/partitioner.cpp:113, is not available here)."""
    blk = (num_vertices + num_parts - 1) // num_parts
    return (np.arange(num_vertices, dtype=np.int64) // blk).astype(np.int32)


def random_parts(num_vertices: int, num_parts: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, num_parts, size=num_vertices).astype(np.int32)


def edge_cut(src, dst, parts) -> float:
    return float(np.mean(parts[src] != parts[dst]))
