"""Last ApplyVertex of a GCN (logits, soft-max, d, d.W^T, AH^T.d) on V rows, timed per variant with CUDA events
on the engine's stream: separate kernels, the fused fp32 kernel, the tcgen05 kernel at 1 / 2 / 4 stages per CTA.

    python tools/last_layer_bench.py --rows 2000000 --dims 16,48,51 --out gpurun_out/last_layer.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))

from dorylus_b200.engine import FORWARD, GCN, Engine  # noqa: E402
from helpers import random_dataset  # noqa: E402

# --op bwd0: ApplyVertex backward of layer 0 (tanh', dW = ah0^T . g) instead
VARIANTS = {"fused-bwd0": dict(), "separate-bwd0": dict(fuse_tanh_bwd=0),
            "separate": dict(fuse_softmax=0), "separate-simt": dict(tensor_cores=0, fuse_softmax=0),
            "fused-simt": dict(fuse_softmax=2), "tc": dict(), "tc-1": dict(tc_stages=1), "tc-2": dict(tc_stages=2),
            "tc-4": dict(tc_stages=4), "tc+tn-tc": dict(tn_small=0), "round-1 tc": dict(tc_small=0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--dims", default="16,48,51")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="separate,fused-simt,tc,tc-2")
    ap.add_argument("--out", default="")
    ap.add_argument("--op", default="last", choices=["last", "bwd0"])
    args = ap.parse_args()
    if args.op == "bwd0":
        return bwd0(args)
    dims = [int(x) for x in args.dims.split(",")]
    L = len(dims) - 1
    V = args.rows
    ds = random_dataset(V=V, E_und=V // 2, dims=dims, seed=3)
    ah = np.random.default_rng(1).standard_normal((V, dims[L - 1]), dtype=np.float32)
    res = dict(rows=V, dims=dims, reps=args.reps, ms={})
    first = None
    for name in args.variants.split(","):
        e = Engine(ds.dims, GCN)
        for k, v in VARIANTS[name].items():
            e.set_option(k, v)
        e.load_partition(ds.images[0])
        e.set_tensor(L - 1, "lab", ds.onehot)
        e.init_weights()
        with e:
            e.set_tensor(L - 1, "ah", ah)
            ch = e.whole_chunk(L - 1, FORWARD)
            for _ in range(3):
                e.applyVertexGCN(ch)
            e.sync()
            e.event_record(0)
            for _ in range(args.reps):
                e.applyVertexGCN(ch)
            e.event_record(1)
            e.sync()
            ms = e.event_elapsed_ms(0, 1) / args.reps
            g = e.get_tensor(L - 1, "grad")
            st = e.stats()
            if first is None:
                first = g
            err = float(np.abs(g - first).max() / max(np.abs(first).max(), 1e-30))
            res["ms"][name] = dict(ms=ms, err_vs_first=err, loss_sum=st["loss_sum"], acc_sum=st["acc_sum"])
            print(name, "%.3f ms" % ms, "err %.2e" % err, flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


def bwd0(args):
    from dorylus_b200.engine import BACKWARD

    dims = [int(x) for x in args.dims.split(",")]
    V = args.rows
    ds = random_dataset(V=V, E_und=V // 2, dims=dims, seed=3)
    rng = np.random.default_rng(1)
    ah = rng.standard_normal((V, dims[0]), dtype=np.float32)
    aTg = rng.standard_normal((V, dims[1]), dtype=np.float32)
    h = np.tanh(rng.standard_normal((V, dims[1]), dtype=np.float32))
    first = None
    for name in ("separate-bwd0", "fused-bwd0"):
        e = Engine(ds.dims, GCN)
        for k, v in VARIANTS[name].items():
            e.set_option(k, v)
        e.load_partition(ds.images[0])
        e.set_tensor(len(dims) - 2, "lab", ds.onehot)
        e.init_weights()
        with e:
            e.set_tensor(0, "ah", ah)
            e.set_tensor(0, "aTg", aTg)
            e.set_tensor(0, "h", h)
            ch = e.whole_chunk(1, BACKWARD)
            for _ in range(3):
                e.applyVertexGCN(ch)
            e.sync()
            e.event_record(0)
            for _ in range(args.reps):
                e.applyVertexGCN(ch)
            e.event_record(1)
            e.sync()
            ms = e.event_elapsed_ms(0, 1) / args.reps
            dw = e.get_weight_grad(0)
            if first is None:
                first = dw
            print(name, "%.3f ms" % ms, "dW err vs separate %.2e" % float(np.abs(dw - first).max() / np.abs(first).max()), flush=True)


if __name__ == "__main__":
    main()
