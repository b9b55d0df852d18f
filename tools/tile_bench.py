#!/usr/bin/env python
"""Aggregation timings of the shared-memory-staged kernel against the gather kernels on one GPU, for a
community-structured synthetic graph (GPU box only).

    python tools/tile_bench.py --name friendster/16 --V 4100522 --E 112500000 --dims 16,48,51 --communities 16384 \
        --variants "tile=0;tile=1;tile=1,tile_rows=256"

Every variant is a ';'-separated set of engine options (k=v,k=v).  Prints one JSON line per variant:
per-aggregation ms, B_alg / t against the measured HBM rate, the plan's coverage.
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
from dorylus_b200.engine import BACKWARD, FORWARD, GCN, Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="friendster/16")
    ap.add_argument("--config", default="", help="a synth.CONFIGS name instead of --V/--E/--dims")
    ap.add_argument("--V", type=int, default=4100522)
    ap.add_argument("--E", type=int, default=112500000)
    ap.add_argument("--dims", default="16,48,51")
    ap.add_argument("--communities", type=int, default=16384)
    ap.add_argument("--locality", type=float, default=0.9)
    ap.add_argument("--sigma", type=float, default=0.9)
    ap.add_argument("--generator", default="streamed", choices=["streamed", "chunglu"])
    ap.add_argument("--variants", default="tile=0;tile=1")
    ap.add_argument("--apply-first", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--epochs", action="store_true", help="also time whole epochs")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if args.config:
        spec = synth.CONFIGS[args.config]
    else:
        spec = synth.GraphSpec(args.name, args.V, args.E, [int(x) for x in args.dims.split(",")], seed=77, sigma=args.sigma,
                               locality=args.locality, communities=args.communities)
    t0 = time.time()
    if args.generator == "streamed" and spec.locality > 0:
        src, dst, _, _ = synth.generate_incident_edges(spec, 0, 1, threads=max(1, (os.cpu_count() or 8) // 2))
    else:
        src, dst = synth.generate_edges(spec)
    V, E, dims = spec.num_vertices, int(src.size), spec.dims
    L = len(dims) - 1
    image = dengine.preprocess_edges(src, dst, np.zeros(V, np.int32), V, 0, 1)
    del src, dst
    print("[tile_bench] %s V=%d E=%d dims=%s built in %.1fs" % (spec.name, V, E, dims, time.time() - t0), file=sys.stderr, flush=True)
    peak = 6550.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    feats = synth.generate_feature_rows(0, V, dims[0], 3)
    onehot = formats.one_hot(synth.generate_label_rows(0, V, dims[-1], 4), dims[-1])
    results = []
    for variant in args.variants.split(";"):
        opts = dict(kv.split("=") for kv in variant.split(",") if kv)
        r = dict(name=spec.name, V=V, E=E, dims=dims, options=opts, apply_first=args.apply_first)
        try:
            with Engine(dims, GCN, flags=_lib.FLAG_APPLY_FIRST if args.apply_first else 0) as e:
                for k, v in opts.items():
                    e.set_option(k, v)
                t0 = time.time()
                e.load_partition(image)
                r["load_s"] = time.time() - t0
                r["tile_fwd"], r["tile_bwd"] = e.tile_info(FORWARD), e.tile_info(BACKWARD)
                e.set_tensor(0, "x", feats)
                e.set_tensor(L - 1, "lab", onehot)
                e.init_weights()
                sched = [e.apply_first(l) for l in range(L)]
                for _ in range(2):
                    e.epoch_async()
                e.sync()
                if args.epochs:
                    e.event_record(0)
                    for _ in range(args.reps):
                        e.epoch_async()
                    e.event_record(1)
                    e.sync()
                    r["epoch_ms"] = e.event_elapsed_ms(0, 1) / args.reps
                cases = [(l, FORWARD) for l in range(L)] + [(l, BACKWARD) for l in range(L - 1, -1, -1) if l > 0 or sched[0]]
                r["layers"] = []
                for layer, d in cases:
                    c = e.whole_chunk(layer, d)
                    F = dims[layer + 1] if sched[layer] else dims[layer]
                    e.aggregate(c)
                    e.event_record(2)
                    for _ in range(args.reps):
                        e.aggregate(c)
                    e.event_record(3)
                    e.sync()
                    ms = e.event_elapsed_ms(2, 3) / args.reps
                    b_alg = 4 * F * V + 4 * F * V + 8 * E + 8 * (V + 1) + 4 * V
                    r["layers"].append(dict(layer=layer, dir="fwd" if d == FORWARD else "bwd", F=F, ms=ms,
                                            edges_per_sec=E / (ms * 1e-3), hbm_frac=b_alg / (ms * 1e-3) / 1e9 / peak,
                                            gathered_tb_per_s=4.0 * F * E / (ms * 1e-3) / 1e12))
        except dengine.DoryError as ex:
            r["error"] = str(ex)
        results.append(r)
        print(json.dumps(r), flush=True)
        brief = {"%s%d F=%d" % (l["dir"], l["layer"], l["F"]): "%.3f ms (%.0f%% HBM)" % (l["ms"], 100 * l["hbm_frac"]) for l in r.get("layers", [])}
        print("[tile_bench] %s -> %s cov %.2f W=%d R=%d epoch %s %s" % (variant, brief, r.get("tile_fwd", {}).get("coverage", 0),
              r.get("tile_fwd", {}).get("window_rows", 0), r.get("tile_fwd", {}).get("tile_rows", 0), r.get("epoch_ms"), r.get("error", "")),
              file=sys.stderr, flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
