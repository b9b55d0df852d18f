#!/usr/bin/env python
"""Per-layer aggregation / epoch timing on one GPU for a synthetic graph of a given shape, with the
HBM-roofline fraction of every aggregation (GPU box only).  Used for the shapes BASELINE.json names
besides Reddit, e.g. one eighth of the Friendster shape (what one of 8 GPUs holds):

    python tools/shape_bench.py --name friendster/8 --V 8201045 --E 225000000 --dims 16,48,51
    python tools/shape_bench.py --name amazon/8     --V 1178761 --E 28949288  --dims 100,64,64,25
    python tools/shape_bench.py --name reddit --gnn GAT            (Reddit GAT 2-layer, configs[2])
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dorylus_b200 import _lib, formats, synth  # noqa: E402
from dorylus_b200 import engine as dengine  # noqa: E402
from dorylus_b200.engine import BACKWARD, FORWARD, GAT, GCN, Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="reddit")
    ap.add_argument("--V", type=int, default=0)
    ap.add_argument("--E", type=int, default=0)
    ap.add_argument("--dims", default="")
    ap.add_argument("--gnn", default="GCN")
    ap.add_argument("--sigma", type=float, default=0.9)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--locality", type=float, default=0.0, help="fraction of edges kept inside a community block")
    ap.add_argument("--communities", type=int, default=1)
    ap.add_argument("--out", default="")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value set before load (repeatable)")
    args = ap.parse_args()
    if args.V:
        spec = synth.GraphSpec(args.name, args.V, args.E, [int(x) for x in args.dims.split(",")], seed=77, sigma=args.sigma,
                               locality=args.locality, communities=args.communities)
    else:
        spec = synth.CONFIGS[args.name]
    t0 = time.time()
    src, dst = synth.generate_edges(spec)
    V, E, dims = spec.num_vertices, int(src.size), spec.dims
    L = len(dims) - 1
    image = dengine.preprocess_edges(src, dst, np.zeros(V, np.int32), V, 0, 1)
    del src, dst
    print("[shape] %s V=%d E=%d dims=%s built in %.1fs" % (spec.name, V, E, dims, time.time() - t0), file=sys.stderr, flush=True)
    gnn = GAT if args.gnn == "GAT" else GCN
    flags = _lib.FLAG_GAT_PREDICT_AH if gnn == GAT else 0
    peak = 6546.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    rng = np.random.default_rng(0)
    out = dict(name=spec.name, gnn=args.gnn, V=V, E=E, dims=dims, hbm_peak_gbs=peak, layers=[])
    out["options"] = args.opt
    with Engine(dims, gnn, flags=flags) as e:
        for kv in args.opt:
            k, v = kv.split("=", 1)
            e.set_option(k, v)
        e.load_partition(image)
        del image
        feats = synth.generate_features(V, dims[0], 3)
        e.set_tensor(0, "x" if gnn == GCN else "h", feats)
        del feats
        e.set_tensor(L - 1, "lab", formats.one_hot(synth.generate_labels(V, dims[-1], 4), dims[-1]))
        e.init_weights()
        for _ in range(2):
            e.epoch_async()
        e.sync()
        e.event_record(0)
        for _ in range(args.reps):
            e.epoch_async()
        e.event_record(1)
        e.sync()
        out["epoch_ms"] = e.event_elapsed_ms(0, 1) / args.reps
        n_agg = (2 * L - 1) if gnn == GCN else 3 * L  # GAT backward walks both adjacencies
        out["aggregated_edges_per_sec"] = n_agg * E / (out["epoch_ms"] * 1e-3)
        cases = ([(l, FORWARD) for l in range(L)] + [(l, BACKWARD) for l in range(L - 1, 0, -1)]) if gnn == GCN else \
                ([(l + 1, FORWARD) for l in range(L)])
        for layer, d in cases:
            c = e.whole_chunk(layer, d)
            F = dims[layer] if gnn == GCN else dims[layer]
            e.aggregate(c)
            e.event_record(2)
            for _ in range(args.reps):
                e.aggregate(c)
            e.event_record(3)
            e.sync()
            ms = e.event_elapsed_ms(2, 3) / args.reps
            b_alg = 4 * F * V + 4 * F * V + 8 * E + 8 * (V + 1) + 4 * V
            out["layers"].append(dict(layer=layer, dir="fwd" if d == FORWARD else "bwd", F=F, ms=ms,
                                      edges_per_sec=E / (ms * 1e-3), algorithmic_bytes=b_alg,
                                      achieved_gbs=b_alg / (ms * 1e-3) / 1e9, hbm_frac=b_alg / (ms * 1e-3) / 1e9 / peak))
        st = e.stats()
        out["loss_sum"], out["acc_sum"] = st["loss_sum"], st["acc_sum"]
    print(json.dumps(out), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
