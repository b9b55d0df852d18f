"""The N>1 host path on CPU: two processes over gloo, one partition each.  The ghost plan
(dorylus_b200/dist.py) is built through torch.distributed exactly as on the GPU box; compute is the
CPU oracle, so this checks the partition / send-list / ghost-slot logic end to end against the
single-partition run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, parts_kind, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from helpers import random_dataset, rel_err
    from dorylus_b200 import dist as ddist
    from oracle.driver import BACKWARD, FORWARD, OracleGCN
    from oracle.pyoracle import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims = [20, 8, 5]
        one = random_dataset(V=260, E_und=1700, dims=dims, P=1, seed=13)
        many = random_dataset(V=260, E_und=1700, dims=dims, P=world, seed=13, parts=parts_kind)
        o = Oracle(build=False)
        o.set_threads(1)
        full = OracleGCN(o, one.graphs, dims)
        full.load_features(one.feats, one.onehot)
        full.epoch()

        g = many.graphs[rank]
        plan = ddist.GhostPlan(g, rank, world)
        plan.exchange_over()
        assert plan.complete()
        # this rank's state machine, ghost rows moved by real messages
        me = OracleGCN(o, [g], dims)
        t = me.saved[0]
        t[0]["x"][:] = many.feats[g.local_to_global]
        t[0]["fg"][:] = many.feats[g.src_ghost_gvid]
        me.aggregate(0, 0, FORWARD)
        me.apply_vertex_forward(0, 0)
        ddist.host_exchange_rows(plan, ddist.FORWARD, t[0]["h"], t[1]["fg"])
        me.aggregate(0, 1, FORWARD)
        errs = {n: rel_err(t[l][n], full.saved[0][l][n][g.local_to_global]) for l, n in ((0, "ah"), (0, "h"), (1, "ah"))}
        t[1]["grad"][:] = full.saved[0][1]["grad"][g.local_to_global]
        ddist.host_exchange_rows(plan, ddist.BACKWARD, t[1]["grad"], t[0]["bg"])
        me.aggregate(0, 1, BACKWARD)
        errs["aTg"] = rel_err(t[0]["aTg"], full.saved[0][0]["aTg"][g.local_to_global])
        q.put((rank, errs, int(g.src_ghost_cnt), int(g.dst_ghost_cnt)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("parts_kind", ["random", "contiguous"])
def test_two_rank_gloo_exchange_matches_single_partition(parts_kind):
    import torch.multiprocessing as mp

    from oracle.pyoracle import Oracle

    Oracle()  # build once in the parent
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, parts_kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs, gs, gd in results:
        assert gs > 0 and gd > 0
        assert max(errs.values()) < 1e-5, (rank, errs)


def test_recv_slots_rejects_foreign_vertex():
    sys.path.insert(0, ROOT)
    from dorylus_b200.dist import recv_slots

    ghosts = np.array([3, 8, 11], np.uint32)
    assert recv_slots(ghosts, np.array([8, 3], np.uint32)).tolist() == [1, 0]
    with pytest.raises(ValueError):
        recv_slots(ghosts, np.array([9], np.uint32))


def test_ghost_slots_from_images_equal_the_exchanged_plan():
    """dory_ghost_slots (host only: the receive plan from two graph.<id>.bin images) equals what
    GhostPlan derives from the id lists the ranks swap -- for every (direction, receiver, sender)."""
    from dorylus_b200 import engine as dengine
    from dorylus_b200.dist import BACKWARD, FORWARD, recv_slots
    from helpers import random_dataset

    ds = random_dataset(V=500, E_und=3000, dims=[8, 4, 3], P=4, seed=9)
    for me, g in enumerate(ds.graphs):
        for peer, gp in enumerate(ds.graphs):
            if peer == me:
                continue
            for d, sends, ghosts in ((FORWARD, gp.fwd_send, g.src_ghost_gvid), (BACKWARD, gp.bwd_send, g.dst_ghost_gvid)):
                want = recv_slots(ghosts, gp.local_to_global[sends[me]])
                got = dengine.ghost_slots(ds.images[me], me, ds.images[peer], d)
                assert np.array_equal(got, want), (me, peer, d)
    # a partition image of a different cut is refused
    other = random_dataset(V=500, E_und=3000, dims=[8, 4, 3], P=2, seed=9)
    with pytest.raises(dengine.DoryError):
        dengine.ghost_slots(ds.images[0], 0, other.images[1], FORWARD)


def _engine_worker(rank, world, port, hostcheck_lib, apply_first, q):
    """One rank of the bench-style start-up on the CPU: gloo for the host plumbing (dist.setup_engine_comm:
    communicator id broadcast, id lists swapped, receive plan installed), the product's engine object on
    the emulated CUDA runtime (tests/hostcheck) for the compute, epochs checked against the oracle."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    import torch.distributed as dist

    from helpers import random_dataset, rel_err
    from dorylus_b200 import _lib
    from dorylus_b200 import dist as ddist
    from dorylus_b200.engine import GCN, Engine
    from oracle.driver import OracleGCN
    from oracle.pyoracle import Oracle

    lib = C.CDLL(hostcheck_lib)
    for name, (res, args) in _lib.SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib._lib = lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dims = [40, 12, 5]
        ds = random_dataset(V=400, E_und=3000, dims=dims, P=world, seed=29)
        o = Oracle(build=False)
        o.set_threads(1)
        orc = OracleGCN(o, ds.graphs, dims)
        orc.load_features(ds.feats, ds.onehot)
        g = ds.graphs[rank]
        with Engine(dims, GCN, node_id=rank, num_nodes=world, device=rank,
                    flags=_lib.FLAG_APPLY_FIRST if apply_first else 0) as e:
            e.load_partition(ds.images[rank])
            e.set_tensor(0, "x", ds.feats[g.local_to_global])
            e.set_tensor(0, "fg", ds.feats[g.src_ghost_gvid])
            e.set_tensor(1, "lab", ds.onehot[g.local_to_global])
            e.init_weights()
            ddist.setup_engine_comm(e, g, rank, world, peer_memory=False)
            errs = {}
            for ep in range(2):
                want = orc.epoch()
                st = e.epoch()
                assert st["acc_sum"] == want["acc"][rank]
                errs["h0.%d" % ep] = rel_err(e.get_tensor(0, "h"), orc.saved[rank][0]["h"])
                errs["aTg0.%d" % ep] = rel_err(e.get_tensor(0, "aTg"), orc.saved[rank][0]["aTg"]) / 2
                for l in range(2):
                    errs["dW%d.%d" % (l, ep)] = rel_err(e.get_weight_grad(l), sum(orc.dW[p][l] for p in range(world))) / 2
                    e.set_weights(l, orc.W[l])
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("apply_first", [False, True], ids=["reference-order", "apply-first"])
def test_two_rank_gloo_engines_on_the_emulated_runtime(apply_first):
    import importlib.util

    import torch.multiprocessing as mp

    from oracle.pyoracle import Oracle

    Oracle()
    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(ROOT, "tests", "hostcheck", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = mod.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, lib, apply_first, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in results:
        assert max(errs.values()) < 1e-5, (rank, errs)
