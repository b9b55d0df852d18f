#!/usr/bin/env python
"""Times the dense apply (vtxNNForward hidden layer: z = ah . W, h = tanh(z)) at Reddit scale on the
fp32 CUDA-core path and on the tcgen05 3xTF32 path, and reports their error against float64 on a
row sample (GPU box only)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dorylus_b200 import engine as dengine  # noqa: E402
from dorylus_b200.engine import FORWARD, GCN, Engine  # noqa: E402


def main():
    V, dims = 232965, [602, 128, 41]
    rng = np.random.default_rng(1)
    # a ring graph: the adjacency is irrelevant here, only V matters
    src = np.arange(V, dtype=np.uint32)
    dst = ((np.arange(V) + 1) % V).astype(np.uint32)
    image = dengine.preprocess_edges(src, dst, np.zeros(V, np.int32), V, 0, 1)
    ah = rng.standard_normal((V, dims[0])).astype(np.float32)
    res = {}
    for tc in (0, 1):
        with Engine(dims, GCN) as e:
            e.set_option("tensor_cores", tc)
            e.load_partition(image)
            e.init_weights()
            e.set_tensor(0, "ah", ah)
            W = e.get_weights(0)
            c = e.whole_chunk(0, FORWARD)
            for _ in range(3):
                e.applyVertexGCN(c)
            e.event_record(0)
            for _ in range(10):
                e.applyVertexGCN(c)
            e.event_record(1)
            e.sync()
            ms = e.event_elapsed_ms(0, 1) / 10
            z = e.get_tensor(0, "z")
            rows = rng.integers(0, V, 2048)
            z64 = ah[rows].astype(np.float64) @ W.astype(np.float64)
            err = float(np.max(np.abs(z[rows] - z64)) / np.max(np.abs(z64)))
            flops = 2.0 * V * dims[0] * dims[1]
            res["tcgen05_3xtf32" if tc else "simt_fp32"] = dict(ms=ms, tflops_effective=flops / ms / 1e9, rel_err_vs_f64=err)
            print(json.dumps({("tc" if tc else "simt"): res["tcgen05_3xtf32" if tc else "simt_fp32"]}), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_bench.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
