#!/usr/bin/env python
"""Tuning sweep of the aggregation kernel shape on the Reddit-shaped graph (GPU box only).

    python tools/spmm_sweep.py [--workload reddit] [--out gpurun_out/spmm_sweep.json]

For every (lanes-per-row LG, float4-per-lane VEC, gather mode U, CTAs/SM OCC, heavy-degree) it times
the layer-0 forward (F=602) and layer-1 forward (F=128) aggregations with CUDA events on the
engine's stream.  Results feed the defaults in csrc/spmm.cu and DESIGN.md §5.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dorylus_b200 import engine as dengine  # noqa: E402
from dorylus_b200 import synth  # noqa: E402
from dorylus_b200.engine import FORWARD, GCN, Engine  # noqa: E402

# (lg, vec, unroll, occ)
SHAPES = {
    0: [(0, 0, 0, 0), (8, 4, 1, 4), (8, 4, 1, 5), (8, 4, 1, 6), (8, 4, 2, 4),
        (8, 2, 1, 4), (8, 2, 1, 6), (8, 2, 1, 8), (8, 2, 2, 4), (8, 2, 2, 5), (8, 2, 2, 6),
        (16, 2, 1, 4), (16, 2, 1, 6), (16, 2, 2, 5), (16, 1, 2, 6), (16, 1, 2, 8), (8, 1, 2, 8), (32, 1, 2, 8),
        (32, 2, 2, 6), (32, 4, 1, 5), (4, 2, 2, 6)],
    1: [(0, 0, 0, 0), (8, 4, 1, 4), (8, 4, 1, 5), (8, 4, 1, 6), (8, 4, 2, 4), (8, 2, 1, 6), (8, 2, 2, 5),
        (8, 2, 2, 6), (16, 2, 1, 6), (16, 2, 2, 5), (32, 1, 2, 8), (4, 2, 2, 6)],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="reddit")
    ap.add_argument("--out", default="gpurun_out/spmm_sweep.json")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--heavy", default="1024")
    ap.add_argument("--src-blocks", default="1")
    args = ap.parse_args()
    spec = synth.CONFIGS[args.workload]
    src, dst = synth.generate_edges(spec)
    parts = np.zeros(spec.num_vertices, np.int32)
    image = dengine.preprocess_edges(src, dst, parts, spec.num_vertices, 0, 1)
    E = int(src.size)
    del src, dst
    feats = synth.generate_features(spec.num_vertices, spec.dims[0], spec.seed + 1)
    rng = np.random.default_rng(0)
    h = rng.standard_normal((spec.num_vertices, spec.dims[1])).astype(np.float32)
    results = []
    combos = [(int(h), int(b)) for h in args.heavy.split(",") for b in args.src_blocks.split(",")]
    for hi, (heavy, nblk) in enumerate(combos):
        with Engine(spec.dims, GCN) as e:
            e.set_option("heavy_degree", heavy)
            e.set_option("src_blocks", nblk)
            e.load_partition(image)
            e.set_tensor(0, "x", feats)
            e.set_tensor(0, "h", h)
            for layer in (0, 1):
                c = e.whole_chunk(layer, FORWARD)
                for lg, vec, un, occ in SHAPES[layer]:
                    if hi > 0 and (lg, vec, un, occ) not in ((0, 0, 0, 0), (8, 4, 1, 4), (8, 4, 2, 4), (16, 2, 1, 4), (8, 2, 2, 5)):
                        continue
                    e.set_option("spmm_lg", lg)
                    e.set_option("spmm_vec", vec)
                    e.set_option("spmm_unroll", un)
                    e.set_option("spmm_occ", occ)
                    try:
                        e.aggregate(c)
                        e.aggregate(c)
                        e.event_record(0)
                        for _ in range(args.reps):
                            e.aggregate(c)
                        e.event_record(1)
                        e.sync()
                        ms = e.event_elapsed_ms(0, 1) / args.reps
                    except dengine.DoryError as ex:
                        ms = None
                        print("skip", layer, lg, vec, un, occ, ex, flush=True)
                    r = dict(layer=layer, F=spec.dims[layer], lg=lg, vec=vec, unroll=un, occ=occ, heavy=heavy,
                             src_blocks=nblk, ms=ms,
                             gedges_per_s=None if ms is None else E / ms / 1e6)
                    results.append(r)
                    print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(workload=args.workload, V=spec.num_vertices, E=E, results=results,
                       when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())), f, indent=1)


if __name__ == "__main__":
    main()
