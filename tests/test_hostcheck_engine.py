"""The engine's HOST logic without a GPU: the product's own engine object (engine_cu.o, unmodified)
linked against a fake CUDA runtime and scalar CPU statements of the kernel launchers
(tests/hostcheck/), driven through the same C ABI and the same Python mirror as on the GPU, and held to
the same oracle comparisons -- the bodies of the `-m gpu` tests are reused as they are.

What this covers: tensor tables and views, padded pitches, operator order and the chunk state machine,
the apply-first schedule (DORY_FLAG_APPLY_FIRST), source-window passes (GCN and, with "gat_windows",
GAT), ghost blocks, error paths.  What it cannot cover: the CUDA kernels themselves -- that is what
`pytest -m gpu` is for."""
import numpy as np
import pytest

import test_gpu_parity as gp
import test_gpu_zzz_apply_first as af
import test_gpu_zz_lambda_golden as lg


@pytest.mark.parametrize("shape", gp.SHAPES, ids=lambda s: "x".join(map(str, s["dims"])))
def test_aggregations_and_epochs_reference_order(hostcheck, oracle, shape):
    gp.test_aggregate_forward_and_backward(oracle, shape)
    if shape in gp.SHAPES[:4]:
        gp.test_epochs_match_oracle(oracle, shape, 0)


def test_operator_level_behaviour_reference_order(hostcheck, oracle, golden):
    gp.test_engine_matches_numpy_gnn_golden(golden)
    lg.test_engine_matches_lambda_ops_golden(golden)
    gp.test_operator_sequence_equals_epoch(oracle)
    gp.test_strict_mask_flag(oracle)
    gp.test_partitions_with_ghosts_single_gpu(oracle)
    gp.test_chunk_subrange_matches_reference_semantics(oracle)
    gp.test_adam_step_matches_oracle_on_identical_gradients(oracle)
    gp.test_shape_and_state_errors()
    gp.test_stream_ordered_stats_readback(oracle)
    gp.test_prefetch_pipeline_semantics(oracle)


@pytest.mark.parametrize("nb", [2, 3, 7])
def test_source_windows(hostcheck, oracle, nb):
    gp.test_source_blocked_aggregation(oracle, nb)


@pytest.mark.parametrize("mode", ["ah", "az"])
def test_gat(hostcheck, oracle, mode):
    gp.test_gat_epoch_matches_oracle(oracle, mode)
    gp.test_gat_quirk_mode_refuses_out_of_bounds_read()


@pytest.mark.parametrize("nb", [2, 5])
def test_gat_source_windows(hostcheck, oracle, nb):
    af.test_gat_source_windows_match_oracle(oracle, nb)


AF_CASES = [
    ([602, 128, 41], 600, 7200, None),
    ([1433, 16, 7], 2708, 5278, None),
    ([100, 64, 64, 25], 2000, 9000, None),
    ([24, 16, 16, 4], 500, 3000, [False, True, False]),
    ([12, 20, 6], 400, 2500, [True, False]),
    ([12, 20, 6], 400, 2500, [False, True]),
    ([16, 48, 51], 1800, 6000, None),
]


@pytest.mark.parametrize("dims,V,E,mask", AF_CASES, ids=lambda v: "x".join(map(str, v)) if isinstance(v, list) else str(v))
def test_apply_first_epochs(hostcheck, oracle, dims, V, E, mask):
    af.test_apply_first_epochs_match_reference_oracle(oracle, dims, V, E, mask)


def test_apply_first_operator_level(hostcheck, oracle):
    af.test_apply_first_tensors_and_errors()
    af.test_apply_first_operator_sequence_equals_epoch()
    af.test_apply_first_two_partitions_on_one_gpu(oracle)


def test_remaining_host_paths(hostcheck, oracle):
    """Row lists (heavy / hub / light, degree classes), tuning options, empty graphs, tiny widths."""
    gp.test_empty_graph_and_tiny_widths(oracle)
    gp.test_aggregate_is_bit_reproducible_and_linear()
    for nb, hub in [(1, 64), (3, 64), (2, 100000)]:
        gp.test_hub_rows_on_thread_block_clusters(oracle, nb, hub)
    for F in (16, 100):
        for light in (1, 2):
            gp.test_light_row_kernels(oracle, F, light)


def _run_ranks(P, body):
    """One host thread per rank (ctypes releases the GIL inside the library, where the emulated
    collectives rendezvous); re-raises the first failure."""
    import threading

    errs = [None] * P

    def run(r):
        try:
            body(r)
        except BaseException as ex:  # noqa: BLE001 -- reported to the main thread
            errs[r] = ex

    th = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(P)]  # daemon: a hung rank must not
    for t in th:                                                                     # keep the test process alive
        t.start()
    for t in th:
        t.join(timeout=180)
    for ex in errs:  # a rank that failed is the cause; the ranks waiting for it are the symptom
        if ex is not None:
            raise ex
    assert not any(t.is_alive() for t in th), "a rank hung in a collective"


@pytest.mark.parametrize("dims,P,mask", [([48, 16, 5], 3, "off"), ([48, 16, 5], 3, None), ([24, 16, 16, 4], 2, [True, False, True]),
                                         ([24, 16, 16, 4], 4, [False, True, False]), ([12, 20, 6], 2, [True, False]),
                                         ([48, 16, 5], 8, None)],
                         ids=["reference-order", "apply-first", "mixed-TFT", "mixed-FTF", "mixed-TF", "apply-first-8-ranks"])
def test_partitions_with_emulated_collectives(hostcheck, oracle, dims, P, mask):
    """P engines, one per partition, each on its own thread, with the ghost exchanges and the dW
    all-reduce carried by the emulated communicator: the whole multi-partition epoch -- receive plan
    from the partition images (dory_ghost_slots), forward and backward exchanges of whatever the
    schedule ships (h / grad, or t / dL/dz for apply-first layers), summed weight gradients, Adam --
    against the oracle's partitioned reference-order run."""
    import threading

    from helpers import random_dataset, rel_err
    from dorylus_b200 import _lib
    from dorylus_b200 import engine as dengine
    from dorylus_b200.engine import GCN, Engine
    from oracle.driver import OracleGCN

    ds = random_dataset(V=500, E_und=4000, dims=dims, P=P, seed=23)
    orc = OracleGCN(oracle, ds.graphs, dims)
    orc.load_features(ds.feats, ds.onehot)
    L = len(dims) - 1
    uid = Engine.comm_unique_id()
    gate = threading.Barrier(P)
    want, checked = {}, []

    def rank(r):
        g = ds.graphs[r]
        flags = _lib.FLAG_APPLY_FIRST if mask is None else 0
        e = Engine(dims, GCN, node_id=r, num_nodes=P, flags=flags)
        if mask not in (None, "off"):
            e.set_option("apply_first_mask", sum(1 << l for l, m in enumerate(mask) if m))
        e.load_partition(ds.images[r])
        with e:
            e.set_tensor(0, "x", ds.feats[g.local_to_global])
            if g.src_ghost_cnt:
                e.set_tensor(0, "fg", ds.feats[g.src_ghost_gvid])
            e.set_tensor(L - 1, "lab", ds.onehot[g.local_to_global])
            e.init_weights()
            e.comm_init(uid)
            for d in (0, 1):
                for q in range(P):
                    if q != r:
                        e.comm_set_recv_slots(d, q, dengine.ghost_slots(ds.images[r], r, ds.images[q], d))
            sched = [e.apply_first(l) for l in range(L)]
            for ep in range(2):
                if gate.wait() == 0:
                    want[ep] = orc.epoch()
                gate.wait()
                st = e.epoch()
                t = orc.saved[r]
                assert st["acc_sum"] == want[ep]["acc"][r]
                for l in range(L - 1):
                    assert rel_err(e.get_tensor(l, "z"), t[l]["z"]) < 1e-5, (r, ep, l, "z")
                    assert rel_err(e.get_tensor(l, "h"), t[l]["h"]) < 1e-5, (r, ep, l, "h")
                    assert rel_err(e.get_tensor(l, "aTg"), t[l]["aTg"]) < 2e-5, (r, ep, l, "aTg")
                for l in range(L):
                    total = sum(orc.dW[p][l] for p in range(P))
                    assert rel_err(e.get_weight_grad(l), total) < 2e-5, (r, ep, l, "dW")
                    if not sched[l]:
                        assert rel_err(e.get_tensor(l, "ah"), t[l]["ah"]) < 1e-5, (r, ep, l, "ah")
                        if l > 0 and g.src_ghost_cnt:
                            assert rel_err(e.get_tensor(l, "fg"), t[l]["fg"]) < 1e-5, (r, ep, l, "fg")
                    assert rel_err(e.get_weights(l), orc.W[l]) < 5e-4, (r, ep, l, "W")
                gate.wait()  # everybody has compared before anybody re-syncs
                for l in range(L):
                    e.set_weights(l, orc.W[l])
                checked.append((r, ep))

    _run_ranks(P, rank)
    assert len(checked) == 2 * P


def test_cpp_driver_host_logic(hostcheck, oracle, monkeypatch):
    """host/dorylus_b200_run.cpp linked against the hostcheck library: the C++ driver's epoch loop, the
    pipeline shell over the real engine (1 and 3 chunks per partition) and --apply-first, against the
    oracle's accuracy / loss per epoch -- the bodies of tests/test_gpu_host_driver.py, plus one run of
    the apply-first flag."""
    import importlib.util
    import os
    import re
    import subprocess

    import test_gpu_host_driver as hd
    from helpers import random_dataset
    from oracle.driver import OracleGCN

    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(hd.ROOT, "tests", "hostcheck", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(hd, "BIN", mod.build_driver())
    hd.test_driver_epochs_match_oracle(oracle)
    for lambdas in (1, 3):
        hd.test_pipeline_mode_matches_oracle(oracle, lambdas)
    ds = random_dataset(V=900, E_und=7000, dims=[50, 16, 6], seed=77)
    cmd = hd.write_dataset(ds)
    r = subprocess.run(cmd + ["--numepochs", "3", "--apply-first", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [(float(m.group(1)), float(m.group(2))) for m in re.finditer(r"Epoch \d+, acc: ([0-9.]+), loss: ([0-9.]+)", r.stdout)]
    orc = OracleGCN(oracle, ds.graphs, ds.dims)
    orc.load_features(ds.feats, ds.onehot)
    val = int(ds.V * 0.1)
    assert len(got) == 3
    for acc, loss in got:
        w = orc.epoch()
        assert abs(acc - w["acc"][0] / val) < 2e-3 and abs(loss - w["loss"][0] / val) < 2e-3
    r = subprocess.run(cmd + ["--apply-first", "1", "--pipeline", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "apply-first" in r.stderr


def test_graft_entry_smoke_host_logic(hostcheck):
    """__graft_entry__.smoke() (what the driver runs on the B200 before the bench), on the emulated runtime."""
    import __graft_entry__ as ge

    ge.smoke()


@pytest.mark.parametrize("P,extra", [(2, []), (3, []), (2, ["--apply-first", "1"])], ids=["2-ranks", "3-ranks", "2-ranks-apply-first"])
def test_cpp_driver_multi_process_host_logic(hostcheck, oracle, P, extra):
    """host/run_onnode.sh with the hostcheck build of the driver: P PROCESSES, each preprocessing and
    loading its partition, deriving the exchange plan from its peers' images, meeting through the
    rendezvous directory (the communicator id), and running epochs with the emulated collectives --
    per-partition accuracy / loss of every epoch against the oracle's partitioned run."""
    import importlib.util
    import os
    import re
    import subprocess

    import test_gpu_host_driver as hd
    from helpers import random_dataset
    from oracle.driver import OracleGCN

    spec = importlib.util.spec_from_file_location("hostcheck_build", os.path.join(hd.ROOT, "tests", "hostcheck", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ds = random_dataset(V=900, E_und=7000, dims=[50, 16, 6], P=P, seed=77)
    cmd = hd.write_dataset(ds)
    env = dict(os.environ, DORY_RUN_BIN=mod.build_driver())
    r = subprocess.run([os.path.join(hd.ROOT, "host", "run_onnode.sh"), str(P)] + cmd[1:] +
                       ["--numepochs", "3", "--exchange", "nccl"] + extra, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = {(int(m.group(1)), int(m.group(2))): (float(m.group(3)), float(m.group(4)))
           for m in re.finditer(r"\[ Node\s+(\d+) \]\s+Epoch (\d+), acc: ([0-9.]+), loss: ([0-9.]+)", r.stdout)}
    assert len(got) == 3 * P, r.stdout
    orc = OracleGCN(oracle, ds.graphs, ds.dims)
    orc.load_features(ds.feats, ds.onehot)
    for ep in (1, 2, 3):
        w = orc.epoch()
        for p, g in enumerate(ds.graphs):
            val = int(g.local_vtx_cnt * 0.1)
            assert abs(got[(p, ep)][0] - w["acc"][p] / val) < 2e-3, (p, ep)
            assert abs(got[(p, ep)][1] - w["loss"][p] / val) < 2e-3, (p, ep)
    for p in range(P):  # every rank preprocessed its own partition: the reference's bytes
        assert open(cmd[2] + "graph.%d.bin" % p, "rb").read() == ds.images[p]


def test_random_models_and_schedules(hostcheck, oracle):
    """16 random models -- 2 to 4 layers, widths from 3 to 130 (every pitch class), any per-layer mix of
    the two schedules, with and without source windows -- two epochs each against the reference-order
    oracle: h, dL/dh and the weight gradients of every layer, and the validation accuracy."""
    from helpers import random_dataset, rel_err
    from dorylus_b200.engine import GCN, Engine
    from oracle.driver import OracleGCN

    rng = np.random.default_rng(0)
    for it in range(16):
        L = int(rng.integers(2, 5))
        dims = [int(rng.choice([3, 7, 16, 17, 33, 48, 64, 100, 130])) for _ in range(L)] + [int(rng.integers(2, 12))]
        mask = [bool(rng.integers(0, 2)) for _ in range(L)]
        V = int(rng.integers(40, 300))
        E = int(rng.integers(V, 8 * V))
        nb = int(rng.choice([0, 0, 2, 3]))
        ds = random_dataset(V=V, E_und=E, dims=dims, seed=100 + it)
        orc = OracleGCN(oracle, ds.graphs, dims)
        orc.load_features(ds.feats, ds.onehot)
        e = Engine(dims, GCN)
        e.set_option("apply_first_mask", sum(1 << l for l, m in enumerate(mask) if m))
        if nb:
            e.set_option("src_blocks", nb)
        e.load_partition(ds.images[0])
        with e:
            e.set_tensor(0, "x", ds.feats)
            e.set_tensor(L - 1, "lab", ds.onehot)
            e.init_weights()
            for ep in range(2):
                want = orc.epoch()
                st = e.epoch()
                case = (it, dims, mask, nb, ep)
                assert st["acc_sum"] == want["acc"][0], case
                for l in range(L - 1):
                    assert rel_err(e.get_tensor(l, "h"), orc.saved[0][l]["h"]) < 1e-5, case
                    assert rel_err(e.get_tensor(l, "aTg"), orc.saved[0][l]["aTg"]) < 2e-5, case
                for l in range(L):
                    assert rel_err(e.get_weight_grad(l), orc.dW[0][l]) < 2e-5, case
                    e.set_weights(l, orc.W[l])


def test_abi_survives_arbitrary_call_sequences(hostcheck):
    """Every entry point answers a bad call with an error code, never with a crash or a stray write
    ("nothing calls exit()/abort()", include/dorylus_b200.h): operators before a partition is loaded,
    chunks with layers / bounds out of range or reversed, every flag combination, unknown tensor names,
    in random order on GCN and GAT engines.  (Run under AddressSanitizer with DORY_HOSTCHECK_LIB; that
    is how the missing bounds check of dory_predict was found.)"""
    from helpers import random_dataset
    from dorylus_b200.engine import GAT, GCN, Chunk, DoryError, Engine

    rng = np.random.default_rng(1)
    errs = ok = 0
    for trial in range(30):
        gnn = GCN if trial % 3 else GAT
        dims = [int(rng.integers(2, 40)) for _ in range(int(rng.integers(3, 5)))]
        ds = random_dataset(V=int(rng.integers(5, 80)), E_und=int(rng.integers(1, 300)), dims=dims, seed=trial)
        try:
            e = Engine(dims, gnn, flags=int(rng.integers(0, 16)))
        except DoryError:
            errs += 1
            continue
        with e:
            if rng.random() < 0.3 and gnn == GCN:
                try:
                    e.set_option("apply_first_mask", int(rng.integers(0, 20)))
                except DoryError:
                    errs += 1
            for fn in (e.aggregate, e.applyVertex, e.scatter, e.applyEdge):  # nothing loaded yet
                with pytest.raises(DoryError):
                    fn(Chunk(0, 0, 0, 1, 0, 0, 0, True))
            e.load_partition(ds.images[0])
            e.set_tensor(0, "x" if gnn == GCN else "h", ds.feats)
            e.set_tensor(len(dims) - 2, "lab", ds.onehot)
            e.init_weights()
            names = ["ah", "z", "h", "grad", "aTg", "t", "g", "u", "fg", "bg", "nope"]
            calls = [lambda: e.forward(int(rng.integers(0, 6))), lambda: e.backward(int(rng.integers(0, 6))), e.epoch,
                     lambda: e.get_tensor(int(rng.integers(0, 5)), names[int(rng.integers(0, len(names)))]),
                     lambda: e.apply_update(int(rng.integers(0, 6)))]
            for _ in range(80):
                if rng.random() < 0.75:
                    c = Chunk(0, 0, int(rng.integers(0, ds.V + 3)), int(rng.integers(0, ds.V + 3)),
                              int(rng.integers(0, len(dims) + 2)), int(rng.integers(0, 2)), 1, bool(rng.integers(0, 2)))
                    call = lambda c=c, fn=[e.aggregate, e.applyVertex, e.scatter, e.applyEdge, e.predictGAT,
                                           e.incLayer][int(rng.integers(0, 6))]: fn(c)
                else:
                    call = calls[int(rng.integers(0, len(calls)))]
                try:
                    call()
                    ok += 1
                except DoryError:
                    errs += 1
    assert ok > 500 and errs > 500


def test_kernel_shape_option_plumbing(hostcheck, oracle):
    """The body of the (4 lanes x 4 float4) shape test: here only its option plumbing and expectations
    are exercised (the scalar statements ignore the kernel shape)."""
    for F in (41, 100):
        af.test_aggregation_shape_4_lanes_by_4_float4(oracle, F)


def test_every_launch_is_within_the_hardware_limits(launchcheck):
    """The product's own launchers (kernel selection, grid arithmetic, cluster attributes of spmm.cu;
    dense.cu; gat.cu) driven by the engine over a sweep of shapes, schedules and options, on a runtime
    that checks every launch against the sm_100 launch limits and executes nothing: one vertex to
    20,000, widths 3 to 1433, hub rows on clusters, every light-row kernel, source windows, GAT."""
    from helpers import random_dataset
    from dorylus_b200 import _lib
    from dorylus_b200.engine import GAT, GCN, Engine
    from test_gpu_parity import HUB

    total = 0
    cases = [
        (GCN, [602, 128, 41], 600, 7200, {}, None), (GCN, [602, 128, 41], 600, 7200, {}, "af"),
        (GCN, [1433, 16, 7], 2708, 5278, {}, "af"), (GCN, [100, 64, 64, 25], 2000, 9000, {"src_blocks": 3}, "af"),
        (GCN, [16, 48, 51], 1800, 6000, {"hub_degree": 64, "heavy_degree": 32}, None),
        (GCN, [3, 5, 2], 1, 1, {}, None), (GCN, [3, 5, 2], 2, 1, {}, "af"),
        (GCN, [130, 17, 9], 20000, 60000, {"spmm_light": 2, "row_order": 2, "locality_block": 100}, None),
        (GCN, [64, 33, 4], 20000, 400000, {"spmm_light": 1, "src_blocks": 5, "hub_degree": 100}, "af"),
        (GCN, [41, 8, 3], 1600, 60000, {"spmm_lg": 4, "spmm_vec": 4, "spmm_light": 1}, None),
        (GCN, [200, 8, 3], 1600, 60000, {"spmm_lg": 32, "spmm_vec": 2, "spmm_unroll": 2, "spmm_occ": 8}, None),
        (GAT, [24, 12, 5], 300, 1500, {}, None), (GAT, [602, 128, 41], 600, 7200, {"gat_windows": 1, "src_blocks": 2}, None),
        # many vertices, tiny widths: the GAT edge-backward workspace has to grow with V (it did not: the
        # launch failed at the Friendster / 8 size, found by running this check at that size)
        (GAT, [8, 4, 3], 600000, 300000, {}, None),
    ]
    for gnn, dims, V, E, opts, sched in cases:
        ds = random_dataset(V=V, E_und=E, dims=dims, seed=7, extra_edges=HUB if V >= 1500 else None)
        flags = (_lib.FLAG_APPLY_FIRST if sched == "af" else 0) | (_lib.FLAG_GAT_PREDICT_AH if gnn == GAT else 0)
        e = Engine(dims, gnn, flags=flags)
        for k, v in opts.items():
            e.set_option(k, v)
        e.load_partition(ds.images[0])
        with e:
            e.set_tensor(0, "x" if gnn == GCN else "h", ds.feats)
            e.set_tensor(len(dims) - 2, "lab", ds.onehot)
            e.init_weights()
            e.epoch()
            e.epoch_async()
            e.stats_enqueue(0)
            e.stats_collect(0)
        n, bad, first = launchcheck()
        assert n > 0 and bad == 0, (dims, V, opts, sched, first)
        total += n
    assert total > 300


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("dims,P,parts,mask", [([48, 16, 5], 3, "random", "off"), ([48, 16, 5], 3, "random", None),
                                               ([48, 16, 5], 4, "contiguous", None), ([24, 16, 16, 4], 2, "random", [True, False, True]),
                                               ([40, 12, 5], 8, "contiguous", "off")],
                         ids=["reference-order", "apply-first", "apply-first-contiguous", "mixed", "reference-order-8-ranks"])
def test_real_communicator_on_emulated_nccl(commcheck, oracle, dims, P, parts, mask, exchange):
    """The product's own communicator (comm.cu host logic: send / receive plans, staging buffers, the
    grouped all-to-all-v with its receive-in-place shortcut for contiguous partitions, and the
    peer-memory path -- send slots, per-peer pointers from the IPC import, the interleaved issue order)
    between P engines on P threads: whole epochs of both schedules against the partitioned oracle,
    the plan computed as the bench does it (send slots = the peer's receive slots for me) and the
    ghost blocks registered through dory_comm_ipc_export / import as dist.setup_peer_memory does."""
    import threading

    from helpers import random_dataset, rel_err
    from dorylus_b200 import _lib
    from dorylus_b200 import engine as dengine
    from dorylus_b200.engine import GCN, Engine
    from oracle.driver import OracleGCN

    ds = random_dataset(V=500, E_und=4000, dims=dims, P=P, seed=23, parts=parts)
    orc = OracleGCN(oracle, ds.graphs, dims)
    orc.load_features(ds.feats, ds.onehot)
    L = len(dims) - 1
    uid = Engine.comm_unique_id()
    gate = threading.Barrier(P)
    want, blobs, checked = {}, [None] * P, []

    def rank(r):
        g = ds.graphs[r]
        e = Engine(dims, GCN, node_id=r, num_nodes=P, flags=_lib.FLAG_APPLY_FIRST if mask is None else 0)
        if mask not in (None, "off"):
            e.set_option("apply_first_mask", sum(1 << l for l, m in enumerate(mask) if m))
        e.load_partition(ds.images[r])
        with e:
            e.set_tensor(0, "x", ds.feats[g.local_to_global])
            if g.src_ghost_cnt:
                e.set_tensor(0, "fg", ds.feats[g.src_ghost_gvid])
            e.set_tensor(L - 1, "lab", ds.onehot[g.local_to_global])
            e.init_weights()
            e.comm_init(uid)
            for d in (0, 1):
                for q in range(P):
                    if q != r:
                        e.comm_set_recv_slots(d, q, dengine.ghost_slots(ds.images[r], r, ds.images[q], d))
                        if exchange == "p2p":  # where MY rows land on q == q's receive slots for me
                            e.comm_set_send_slots(d, q, dengine.ghost_slots(ds.images[q], q, ds.images[r], d))
            if exchange == "p2p":
                blobs[r] = {key: e.comm_ipc_export(*key) for key in e.ghost_tensors()}
                gate.wait()
                for q in range(P):
                    if q != r:
                        for (layer, name), blob in blobs[q].items():
                            e.comm_ipc_import(layer, name, q, blob)
            sched = [e.apply_first(l) for l in range(L)]
            if not sched[0] and g.src_ghost_cnt:
                # the layer-0 input exchange bench.py uses (no reference counterpart): overwrite the ghost
                # features with noise, ship the owned rows of x, and find the owners' rows there, to the bit
                e.set_tensor(0, "fg", np.full((g.src_ghost_cnt, dims[0]), 7.0, np.float32))
            gate.wait()
            if not sched[0]:
                from dorylus_b200.engine import FORWARD

                e.scatter(e.whole_chunk(0, FORWARD))
                if g.src_ghost_cnt:
                    assert np.array_equal(e.get_tensor(0, "fg"), ds.feats[g.src_ghost_gvid])
            for ep in range(2):
                if gate.wait() == 0:
                    want[ep] = orc.epoch()
                gate.wait()
                st = e.epoch()
                t = orc.saved[r]
                assert st["acc_sum"] == want[ep]["acc"][r]
                for l in range(L - 1):
                    assert rel_err(e.get_tensor(l, "h"), t[l]["h"]) < 1e-5, (r, ep, l, "h")
                    assert rel_err(e.get_tensor(l, "aTg"), t[l]["aTg"]) < 2e-5, (r, ep, l, "aTg")
                for l in range(L):
                    assert rel_err(e.get_weight_grad(l), sum(orc.dW[p][l] for p in range(P))) < 2e-5, (r, ep, l, "dW")
                    if not sched[l] and l > 0 and g.src_ghost_cnt:
                        assert rel_err(e.get_tensor(l, "fg"), t[l]["fg"]) < 1e-5, (r, ep, l, "fg")
                gate.wait()
                for l in range(L):
                    e.set_weights(l, orc.W[l])
                checked.append((r, ep))
            gate.wait()  # nobody frees memory a peer may still store into

    _run_ranks(P, rank)
    assert len(checked) == 2 * P


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("windows", [0, 1], ids=["no-windows", "source-windows"])
def test_gat_partitions_on_the_real_communicator(commcheck, oracle, exchange, windows):
    """GAT on several partitions (z forward, grad backward through fg_z / bg_d) -- a combination no GPU
    suite covers yet -- on the real communicator over the emulated NCCL / IPC, against OracleGAT's
    partitioned epoch."""
    import threading

    from helpers import random_dataset, rel_err
    from dorylus_b200 import _lib
    from dorylus_b200 import engine as dengine
    from dorylus_b200.engine import GAT, Engine
    from oracle.driver import OracleGAT

    dims, P = [24, 12, 5], 3
    ds = random_dataset(V=400, E_und=2500, dims=dims, P=P, seed=71)
    orc = OracleGAT(oracle, ds.graphs, dims, predict_from="ah")
    orc.load_features(ds.feats, ds.onehot)
    orc.epoch()
    uid = Engine.comm_unique_id()
    gate = threading.Barrier(P)
    blobs, done = [None] * P, []

    def rank(r):
        g = ds.graphs[r]
        e = Engine(dims, GAT, node_id=r, num_nodes=P, flags=_lib.FLAG_GAT_PREDICT_AH)
        if windows:
            e.set_option("gat_windows", 1)
            e.set_option("src_blocks", 3)
        e.load_partition(ds.images[r])
        with e:
            e.set_tensor(0, "h", ds.feats[g.local_to_global])
            e.set_tensor(1, "lab", ds.onehot[g.local_to_global])
            e.init_weights()
            for l in range(2):
                e.set_weights(l, orc.a[l], "a_i")
            e.comm_init(uid)
            for d in (0, 1):
                for q in range(P):
                    if q != r:
                        e.comm_set_recv_slots(d, q, dengine.ghost_slots(ds.images[r], r, ds.images[q], d))
                        if exchange == "p2p":
                            e.comm_set_send_slots(d, q, dengine.ghost_slots(ds.images[q], q, ds.images[r], d))
            if exchange == "p2p":
                blobs[r] = {key: e.comm_ipc_export(*key) for key in e.ghost_tensors()}
                gate.wait()
                for q in range(P):
                    if q != r:
                        for (layer, name), blob in blobs[q].items():
                            e.comm_ipc_import(layer, name, q, blob)
            gate.wait()
            e.epoch()
            t = orc.saved[r]
            for l in range(2):
                for name in ("z", "ah", "grad", "aTg"):
                    assert rel_err(e.get_tensor(l, name), t[l][name]) < 1e-5, (r, l, name)
                if g.src_ghost_cnt:
                    assert rel_err(e.get_tensor(l, "fg_z"), t[l]["fg_z"]) < 1e-5, (r, l, "fg_z")
                if g.dst_ghost_cnt:
                    assert rel_err(e.get_tensor(l, "bg_d"), t[l]["bg_d"]) < 1e-5, (r, l, "bg_d")
                # the weight gradient every rank holds after the epoch is the all-reduced sum
                assert rel_err(e.get_weight_grad(l), sum(orc.dW[p][l] for p in range(P))) < 1e-5, (r, l, "dW")
            done.append(r)
            gate.wait()

    _run_ranks(P, rank)
    assert len(done) == P
