#!/usr/bin/env python
"""Runs a few layer-0 / layer-1 forward aggregations on the Reddit-shaped graph with a given kernel
shape, for capture under ncu (GPU box only):

    ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 4 -c 2 \
        -o gpurun_out/prof python tools/profile_spmm.py --cfg 8,2,2,5 --layer 0
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dorylus_b200 import engine as dengine  # noqa: E402
from dorylus_b200 import synth  # noqa: E402
from dorylus_b200.engine import FORWARD, GCN, Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="0,0,0,0", help="lg,vec,unroll,occ")
    ap.add_argument("--layer", type=int, default=0)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--option", action="append", default=[], help="extra key=value engine options")
    args = ap.parse_args()
    spec = synth.CONFIGS["reddit"]
    src, dst = synth.generate_edges(spec)
    image = dengine.preprocess_edges(src, dst, np.zeros(spec.num_vertices, np.int32), spec.num_vertices, 0, 1)
    del src, dst
    with Engine(spec.dims, GCN) as e:
        for kv in args.option:
            k, v = kv.split("=")
            e.set_option(k, v)
        e.load_partition(image)
        if args.layer == 0:
            e.set_tensor(0, "x", synth.generate_features(spec.num_vertices, spec.dims[0], spec.seed + 1))
        else:
            e.set_tensor(0, "h", np.random.default_rng(0).standard_normal((spec.num_vertices, spec.dims[1])).astype(np.float32))
        lg, vec, un, occ = (int(x) for x in args.cfg.split(","))
        for k, v in (("spmm_lg", lg), ("spmm_vec", vec), ("spmm_unroll", un), ("spmm_occ", occ)):
            e.set_option(k, v)
        c = e.whole_chunk(args.layer, FORWARD)
        e.event_record(0)
        for _ in range(args.reps):
            e.aggregate(c)
        e.event_record(1)
        e.sync()
        print("cfg", args.cfg, "layer", args.layer, "ms/aggregate", e.event_elapsed_ms(0, 1) / args.reps)


if __name__ == "__main__":
    main()
