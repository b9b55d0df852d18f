#!/usr/bin/env python
"""Times dory_preprocess_dir against the reference's own DataLoader::preprocess (compiled in
oracle/_ref from graph/dataloader.cpp:225-330) on the same graph.bsnap.edges / .parts, and checks
that the two graph.<id>.bin files are byte-identical.  CPU only; needs /root/reference (or the
prebuilt oracle/_ref).  Lives under tests/ because it executes the oracle side as the baseline.

    python tests/bench_preprocess_vs_reference.py [--V 100000 --E 20000000] [--out profiles/x.json]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dorylus_b200 import engine as dengine  # noqa: E402
from dorylus_b200 import formats, synth  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--V", type=int, default=100000)
    ap.add_argument("--E", type=int, default=20_000_000)
    ap.add_argument("--parts", default="1,4")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    ref = Ref()
    spec = synth.GraphSpec("pp", args.V, args.E, [16, 8, 4], seed=5, sigma=1.0)
    src, dst = synth.generate_edges(spec)
    out = []
    for P in (int(x) for x in args.parts.split(",")):
        parts = synth.contiguous_parts(args.V, P) if P > 1 else np.zeros(args.V, np.int32)
        d = tempfile.mkdtemp() + "/"
        try:
            formats.write_bsnap_edges(d + "graph.bsnap.edges", args.V, src, dst)
            formats.write_parts(d + "graph.bsnap.parts", parts)
            t0 = time.time()
            f = ref.preprocess(d, 0, P, False)
            t_ref = time.time() - t0
            ref_bytes = open(f, "rb").read()
            os.remove(f)
            t0 = time.time()
            dengine.preprocess_dir(d, 0, P, False)
            t_ours = time.time() - t0
            ours = open(d + "graph.0.bin", "rb").read()
        finally:
            shutil.rmtree(d, ignore_errors=True)
        rec = dict(V=args.V, E=int(src.size), partitions=P, partition=0, reference_s=round(t_ref, 2), ours_s=round(t_ours, 2),
                   speedup=round(t_ref / t_ours, 1), byte_identical=ours == ref_bytes, image_mb=round(len(ours) / 1e6, 1),
                   host_cores=os.cpu_count(), reference="DataLoader::preprocess, 1 thread (as written)",
                   ours="dory_preprocess_dir, all host cores")
        print(json.dumps(rec), flush=True)
        out.append(rec)
    if args.out:
        with open(args.out, "w") as fo:
            json.dump(out, fo, indent=1)


if __name__ == "__main__":
    main()
