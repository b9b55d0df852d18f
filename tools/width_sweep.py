#!/usr/bin/env python
"""Aggregation-kernel shape sweep at a given row width on the Reddit-shaped graph (GPU box only).

    python tools/width_sweep.py --widths 41,64,128 [--out gpurun_out/width_sweep.json]

The apply-first schedule moves the Reddit aggregations from 602 / 128 floats per edge to
128 / 41: the 41-float rows (pitch 64, 11 float4) run at ~10 TB/s of gathered bytes where the 128-float
rows reach 18, i.e. they are bound by instructions per byte, not by L2.  This sweeps lanes per row x
float4 per lane x unroll x CTAs/SM (incl. the 4 x 4 shape: 8 edges per gather instruction) for such
widths; an engine with dims [F, 8, 3] aggregates F-wide rows at layer 0.
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

# (lg, vec, unroll, occ); (0, 0, 0, 0) = the defaults
SHAPES = [(0, 0, 0, 0), (4, 4, 1, 4), (4, 4, 1, 6), (4, 4, 2, 4), (4, 4, 2, 6), (8, 2, 1, 4), (8, 2, 1, 6), (8, 2, 2, 6),
          (8, 2, 1, 8), (16, 1, 1, 6), (16, 1, 2, 8), (4, 2, 2, 6), (8, 4, 1, 4)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="reddit")
    ap.add_argument("--widths", default="41,64,128")
    ap.add_argument("--out", default="gpurun_out/width_sweep.json")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    spec = synth.CONFIGS[args.workload]
    src, dst = synth.generate_edges(spec)
    V = spec.num_vertices
    image = dengine.preprocess_edges(src, dst, np.zeros(V, np.int32), V, 0, 1)
    E = int(src.size)
    del src, dst
    rng = np.random.default_rng(0)
    results = []
    for F in (int(x) for x in args.widths.split(",")):
        x = rng.standard_normal((V, F)).astype(np.float32)
        with Engine([F, 8, 3], GCN) as e:
            e.load_partition(image)
            e.set_tensor(0, "x", x)
            c = e.whole_chunk(0, FORWARD)
            for lg, vec, un, occ in SHAPES:
                for k, v in (("spmm_lg", lg), ("spmm_vec", vec), ("spmm_unroll", un), ("spmm_occ", occ)):
                    e.set_option(k, v)
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
                    print("skip", F, lg, vec, un, occ, ex, flush=True)
                r = dict(F=F, lg=lg, vec=vec, unroll=un, occ=occ, ms=ms,
                         gathered_tb_per_s=None if ms is None else 4.0 * ((F + 3) // 4 * 4) * E / (ms * 1e-3) / 1e12)
                results.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(workload=args.workload, V=V, E=E, results=results,
                       when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())), f, indent=1)


if __name__ == "__main__":
    main()
