#!/usr/bin/env python
"""How often does a tile of destination rows see the same source row?  (CPU only.)

    python tools/reuse_stats.py [--workload reddit] [--tiles 8,64,512] [--locality 0.8 --communities 41]

For a tile of T destination rows (consecutive in the engine's issue order: degree-descending for a
graph without locality, vertex order with it) the reuse factor is  edges(tile) / distinct sources(tile):
what an aggregation kernel that keeps a tile's accumulators on chip and stages every source row once
could save in L2 -> SM traffic over the per-edge gather.
"""
import argparse
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dorylus_b200 import synth  # noqa: E402


def rmat_edges(n_und: int, scale: int, seed: int, a=0.57, b=0.19, c=0.19):
    """R-MAT: every edge picks one quadrant per bit level (a: top-left, b: top-right, c: bottom-left, d: the rest)."""
    rng = np.random.default_rng(seed)
    u = np.zeros(n_und, np.int64)
    v = np.zeros(n_und, np.int64)
    for _ in range(scale):
        r = rng.random(n_und, dtype=np.float32)
        u = (u << 1) | (r >= a + b)                        # bottom half: c or d
        v = (v << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c))  # right half: b or d
    keep = u != v
    lo, hi = np.minimum(u[keep], v[keep]), np.maximum(u[keep], v[keep])
    key = np.unique(lo * (1 << scale) + hi)
    lo, hi = (key >> scale).astype(np.uint32), (key & ((1 << scale) - 1)).astype(np.uint32)
    return np.concatenate([lo, hi]), np.concatenate([hi, lo]), n_und


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="reddit")
    ap.add_argument("--tiles", default="8,64,512,4096")
    ap.add_argument("--locality", type=float, default=None)
    ap.add_argument("--communities", type=int, default=None)
    ap.add_argument("--out", default="")
    ap.add_argument("--rmat", action="store_true",
                    help="the generator SURVEY.md 8d names instead of synth.py's: R-MAT (a, b, c = .57, .19, .19) over 2^18 "
                         "vertices, the workload's edge count drawn, symmetrised, duplicates and self loops removed")
    args = ap.parse_args()
    spec = synth.CONFIGS[args.workload]
    if args.locality is not None:
        spec = dataclasses.replace(spec, locality=args.locality, communities=args.communities or 41)
    if args.rmat:
        src, dst, drawn = rmat_edges(spec.num_edges // 2, 18, seed=11)
        spec = dataclasses.replace(spec, name=spec.name + "-rmat", num_vertices=1 << 18)
        print("R-MAT: %d undirected edges drawn, %d distinct after de-duplication (%.1f %% kept)"
              % (drawn, src.size // 2, 100.0 * (src.size // 2) / drawn), file=sys.stderr)
    else:
        src, dst = synth.generate_edges(spec)
    V, E = spec.num_vertices, src.size
    deg = np.bincount(dst, minlength=V)
    if spec.locality > 0:
        rank = np.arange(V, dtype=np.int64)  # vertex order keeps communities together
        order_name = "vertex order"
    else:
        rank = np.empty(V, dtype=np.int64)
        rank[np.argsort(-deg, kind="stable")] = np.arange(V)
        order_name = "degree-descending"
    out = dict(workload=spec.name, V=V, E=int(E), locality=spec.locality, communities=spec.communities,
               order=order_name, tiles=[])
    r = rank[dst]
    for T in (int(t) for t in args.tiles.split(",")):
        tile = r // T
        key = tile.astype(np.uint64) * np.uint64(V) + src.astype(np.uint64)
        distinct = np.unique(key).size
        ntiles = int(tile.max()) + 1
        # per-tile reuse, edge-weighted: sum over tiles of edges_t * (edges_t / distinct_t) / E
        uk = np.unique(key)
        d_t = np.bincount((uk // np.uint64(V)).astype(np.int64), minlength=ntiles)
        e_t = np.bincount(tile, minlength=ntiles)
        with np.errstate(divide="ignore", invalid="ignore"):
            per = np.where(d_t > 0, e_t / np.maximum(d_t, 1), 0.0)
        row = dict(T=T, tiles=ntiles, edges_per_distinct_source=float(E / distinct),
                   edge_weighted_mean_reuse=float((per * e_t).sum() / E),
                   share_of_edges_in_tiles_with_reuse_ge_2=float(e_t[per >= 2].sum() / E),
                   accumulator_bytes_128_float_slab=T * 512)
        out["tiles"].append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
