#!/usr/bin/env python
"""Where one synchronous GCN epoch spends its time, operator by operator, at N ranks (GPU box only).

    python tools/op_breakdown.py                                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/op_breakdown.py [--workload reddit] [--exchange p2p|nccl]

The epoch is driven operator by operator through the C ABI (the same calls dory_epoch makes, in the
same order) with a CUDA event on the engine's stream between consecutive operators; each figure is
the mean over --reps epochs, max over ranks.  The whole-epoch time measured the same way is printed
beside the sum so that the cost of the extra events is visible.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="reddit")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--out", default="")
    ap.add_argument("--emulate-parts", type=int, default=0,
                    help="single GPU: run partition 0 of P (ghost rows filled with noise, no exchange) -- the "
                         "per-rank kernel shapes of a P-GPU run, e.g. under ncu")
    ap.add_argument("--sustain", type=int, default=150, help="epochs of the sustained (clock-sampled) run")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--apply-first", action="store_true", help="DORY_FLAG_APPLY_FIRST (not with --emulate-parts)")
    args = ap.parse_args()
    assert not (args.apply_first and args.emulate_parts), "--emulate-parts fills the reference order's ghost blocks only"
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    import bench
    from dorylus_b200 import _lib as dlib
    from dorylus_b200 import dist as ddist
    from dorylus_b200 import formats
    from dorylus_b200.engine import BACKWARD, FORWARD, GCN, Engine

    emu = args.emulate_parts if world == 1 else 0
    streamed = bench.ARMS.get(args.workload + "_gcn", {}).get("streamed", False)
    wl = bench.build_workload(args.workload, emu or world, rank, streamed=streamed, dist=dist)
    spec, image, graph, n_edges, cut = wl.spec, wl.image, wl.graph, wl.n_edges, wl.cut
    dims = spec.dims
    L = len(dims) - 1
    e = Engine(dims, GCN, node_id=rank, num_nodes=emu or world, device=local,
               flags=dlib.FLAG_APPLY_FIRST if args.apply_first else 0)
    for kv in args.opt:
        k, v = kv.split("=", 1)
        e.set_option(k, v)
    e.load_partition(image)
    e.set_tensor(0, "x", wl.x_loc)
    if emu:
        rng = np.random.default_rng(1)
        e.set_tensor(0, "fg", rng.standard_normal((graph.src_ghost_cnt, dims[0])).astype(np.float32))
        e.set_tensor(1, "fg", rng.standard_normal((graph.src_ghost_cnt, dims[1])).astype(np.float32))
        e.set_tensor(0, "bg", rng.standard_normal((graph.dst_ghost_cnt, dims[1])).astype(np.float32))
    e.set_tensor(L - 1, "lab", wl.onehot)
    e.init_weights()
    sched = [e.apply_first(l) for l in range(L)]
    if world > 1:
        ddist.setup_engine_comm(e, graph, rank, world, peer_memory=args.exchange == "p2p")
        if not sched[0]:  # an apply-first layer 0 gathers t = x . W: no ghost rows of x
            e.scatter(e.whole_chunk(0, FORWARD))

    # the operator sequence of dory_epoch (csrc/engine.cu: dory_forward / dory_backward) for GCN
    ops = []
    for l in range(L):
        c = e.whole_chunk(l, FORWARD)
        if sched[l]:  # AV -> SC -> GA (+ activation)
            ops.append(("AV fwd L%d (t = in.W)" % l, e.applyVertex, c))
            if not emu:
                ops.append(("SC fwd L%d (t)" % l, e.scatter, c))
            ops.append(("GA fwd L%d (F=%d) + act" % (l, dims[l + 1]), e.aggregate, c))
        else:
            ops.append(("GA fwd L%d (F=%d)" % (l, dims[l]), e.aggregate, c))
            ops.append(("AV fwd L%d" % l, e.applyVertex, c))
        n = e.incLayer(c)
        if not emu and not (n.dir == FORWARD and sched[n.layer]):
            ops.append(("SC %s L%d" % ("fwd" if n.dir == FORWARD else "bwd", n.layer), e.scatter, n))
    for l in list(range(L - 1, 0, -1)) + ([0] if sched[0] else []):
        c = e.whole_chunk(l, BACKWARD)
        ops.append(("GA bwd L%d (F=%d)" % (l, dims[l + 1] if sched[l] else dims[l]), e.aggregate, c))
        ops.append(("AV bwd L%d" % (l if sched[l] else l - 1), e.applyVertex, c))
        if l == 0:
            continue
        n = e.incLayer(c)
        if (n.layer != 0 or sched[0]) and not emu:
            ops.append(("SC bwd L%d" % n.layer, e.scatter, n))
    for l in range(L - 1, -1, -1):
        ops.append(("update W%d" % l, lambda layer, _l=l: e.apply_update(_l), l))
    assert len(ops) + 1 <= 60

    def run_epoch(record):
        for i, (_, fn, arg) in enumerate(ops):
            if record:
                e.event_record(i)
            fn(arg)
        if record:
            e.event_record(len(ops))

    def barrier():
        e.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(2):
        run_epoch(False)
    barrier()
    acc = np.zeros(len(ops))
    tot = 0.0
    for _ in range(args.reps):
        run_epoch(True)
        e.sync()
        acc += [e.event_elapsed_ms(i, i + 1) for i in range(len(ops))]
        tot += e.event_elapsed_ms(0, len(ops))
    acc /= args.reps
    tot /= args.reps
    # whole epochs back to back (what bench.py times)
    barrier()
    e.event_record(60)
    for _ in range(0 if emu else args.reps):
        e.epoch_async()
    e.event_record(61)
    barrier()
    plain = e.event_elapsed_ms(60, 61) / args.reps
    if emu:
        plain = tot
    vec = np.concatenate([acc, [tot, plain]])
    # ghost exchange alone, store-kernel variants (option "p2p_rows"), back to back: ranks stay in
    # lockstep through the barriers, so this is the cost of the exchange itself
    sc = {}
    if world > 1 and args.exchange == "p2p":
        c1 = e.whole_chunk(1, FORWARD)
        for label, rows, elide in (("fence+2 barriers", 9, 0), ("1 row/warp, 2 barriers", 1, 0), ("1 row/warp", 1, 1),
                                   ("2 rows/warp", 2, 1), ("4 rows/warp", 0, 1)):
            e.set_option("p2p_rows", rows)
            e.set_option("p2p_elide_barrier", elide)
            for _ in range(3):
                e.scatter(c1)
            barrier()
            e.event_record(58)
            for _ in range(20):
                e.scatter(c1)
            e.event_record(59)
            barrier()
            sc[label] = e.event_elapsed_ms(58, 59) / 20
        e.set_option("p2p_rows", 0)
        e.set_option("p2p_elide_barrier", 1)
    vec = np.concatenate([vec, list(sc.values())])
    # sustained run with every GPU's clocks / power sampled while it lasts (rank 0 samples all GPUs)
    clk = None
    sampler = None
    if rank == 0:
        import subprocess
        import tempfile
        fd, spath = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        try:
            sampler = subprocess.Popen(["nvidia-smi", "--query-gpu=index,clocks.sm,power.draw,temperature.gpu,"
                                        "clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,"
                                        "clocks_event_reasons.sw_thermal_slowdown",
                                        "--format=csv,noheader,nounits", "-lms", "50"],
                                       stdout=open(spath, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            sampler = None
    barrier()
    e.event_record(56)
    for _ in range(0 if emu else args.sustain):
        e.epoch_async()
    e.event_record(57)
    barrier()
    sustained = e.event_elapsed_ms(56, 57) / max(args.sustain, 1)
    vec = np.concatenate([vec, [sustained]])
    if sampler is not None:
        sampler.terminate()
        sampler.wait()
        per = {}
        for line in open(spath):
            c = [x.strip() for x in line.split(",")]
            try:
                per.setdefault(int(c[0]), []).append((float(c[1]), float(c[2]), float(c[3]), c[4:]))
            except (ValueError, IndexError):
                continue
        os.unlink(spath)
        clk = {g: dict(sm_mhz_median=float(np.median([r[0] for r in rows])), sm_mhz_min=min(r[0] for r in rows),
                       power_w_max=max(r[1] for r in rows), temp_c_max=max(r[2] for r in rows), samples=len(rows),
                       power_cap_samples=sum(1 for r in rows if r[3] and r[3][0].lower().startswith("active")))
               for g, rows in sorted(per.items())}
    if dist is not None:
        t = torch.tensor(vec, dtype=torch.float64, device="cuda")
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per_rank = np.stack([a.cpu().numpy() for a in allv])
    else:
        per_rank = vec[None, :]
    vmax, vmin = per_rank.max(0), per_rank.min(0)
    n_ops = len(ops)
    if rank == 0:
        out = dict(workload=args.workload, n_gpus=world, exchange=args.exchange if world > 1 else None,
                   V=spec.num_vertices, E=n_edges, dims=dims, edge_cut=cut, options=args.opt,
                   ops=[dict(op=name, ms_max=float(vmax[i]), ms_min=float(vmin[i]),
                             ms_per_rank=[float("%.4f" % x) for x in per_rank[:, i]]) for i, (name, _, _) in enumerate(ops)],
                   epoch_ms_with_events=float(vmax[n_ops]), epoch_ms=float(vmax[n_ops + 1]),
                   scatter_L1_fwd_ms={k: float(vmax[n_ops + 2 + i]) for i, k in enumerate(sc)},
                   sustained_epoch_ms=None if emu else float(vmax[-1]), sustained_epochs=0 if emu else args.sustain,
                   gpu_clocks_sustained=clk)
        print(json.dumps(out), flush=True)
        for o in out["ops"]:
            print("  %-22s %8.3f ms  (min over ranks %8.3f)  %s" % (o["op"], o["ms_max"], o["ms_min"], o["ms_per_rank"]),
                  file=sys.stderr)
        print("  %-22s %8.3f ms   back-to-back epochs: %.3f ms" % ("sum of operators", float(vmax[:n_ops].sum()), plain),
              file=sys.stderr)
        for k, v in out["scatter_L1_fwd_ms"].items():
            print("  scatter L1 fwd [%s]: %.3f ms" % (k, v), file=sys.stderr)
        if not emu:  # an emulated partition has no peers to run whole epochs with
            print("  sustained (%d epochs): %.3f ms/epoch; clocks %s" % (args.sustain, out["sustained_epoch_ms"], clk), file=sys.stderr)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(out, f, indent=1)
    e.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
