"""The C++ host driver (host/dorylus_b200_run.cpp) over the C ABI: reads the reference's dataset
directory layout (graph.bsnap.edges/.parts, features.bsnap, labels.bsnap, layer config), preprocesses
the partition like Engine::init does, runs epochs, prints the weight server's "Epoch n, acc, loss"
lines -- compared with the CPU oracle's epochs."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from helpers import random_dataset
from dorylus_b200 import formats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "dorylus_b200_run")


def test_driver_is_built():
    assert os.path.exists(BIN), "run `python -m dorylus_b200.build`"


def test_pipeline_shell_unit_checks(tmp_path):
    """host/saga_pipeline.hpp with recording stubs (no GPU): queue priority == Chunk::operator<,
    operator order of GCN (2 and 3 layers, 2 chunks) and GAT epochs, barriers, early stop."""
    exe = str(tmp_path / "test_saga_pipeline")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "host", "test_saga_pipeline.cpp"),
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr


def write_dataset(ds):
    root = tempfile.mkdtemp()
    d = os.path.join(root, "parts_%d" % ds.P) + "/"
    os.makedirs(d)
    formats.write_bsnap_edges(d + "graph.bsnap.edges", ds.V, ds.src, ds.dst)
    formats.write_parts(d + "graph.bsnap.parts", ds.parts.astype(np.int32))
    formats.write_features(os.path.join(root, "features.bsnap"), ds.feats)
    formats.write_labels(os.path.join(root, "labels.bsnap"), ds.labels, ds.dims[-1])
    formats.write_layer_config(os.path.join(root, "layers.config"), ds.dims)
    return [BIN, "--datasetdir", d, "--featuresfile", os.path.join(root, "features.bsnap"),
            "--labelsfile", os.path.join(root, "labels.bsnap"), "--layerfile", os.path.join(root, "layers.config")]


def test_driver_dry_run_reads_the_dataset_directory():
    """--dry-run 1: the host-side half of the driver without a GPU -- preprocess graph.0.bin, read
    features / labels through dory_read_features / dory_read_labels, write the feats<F0>.0.bin cache."""
    ds = random_dataset(V=300, E_und=1500, dims=[11, 6, 4], seed=79)
    cmd = write_dataset(ds)
    d = cmd[2]
    r = subprocess.run(cmd + ["--dry-run", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    m = re.search(r"dry run: V (\d+) ghosts (\d+) F0 (\d+) classes (\d+) feature_sum ([-0-9.]+) label_sum ([0-9.]+)", r.stdout)
    assert m, r.stdout
    assert (int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))) == (300, 0, 11, 4)
    assert abs(float(m.group(5)) - float(ds.feats.astype(np.float64).sum())) < 1e-3
    assert float(m.group(6)) == float(ds.labels.sum())
    assert open(d + "graph.0.bin", "rb").read() == ds.images[0]
    assert np.array_equal(np.fromfile(d + "feats11.0.bin", dtype=np.float32), ds.feats.ravel())
    # a second run takes the cache (the features file may be gone, engine/utils.cpp:488-501)
    r = subprocess.run([c if c != cmd[4] else cmd[4] + ".gone" for c in cmd] + ["--dry-run", "1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "dry run: V 300" in r.stdout, r.stderr


def test_driver_dry_run_multi_partition_plan():
    """--numnodes 3 --nodeid 1 --dry-run 1: the driver preprocesses its own partition (and, having no
    peers in a dry run, theirs), reads its local AND ghost feature rows, and derives the receive / send
    plan of every exchange from the partition images (dory_ghost_slots) -- compared with the plan the
    Python mirror builds from the id lists the ranks swap."""
    from dorylus_b200.dist import recv_slots

    ds = random_dataset(V=400, E_und=2500, dims=[9, 6, 4], P=3, seed=83)
    cmd = write_dataset(ds)
    d, me = cmd[2], 1
    r = subprocess.run(cmd + ["--dry-run", "1", "--numnodes", "3", "--nodeid", str(me)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    g = ds.graphs[me]
    m = re.search(r"dry run: V (\d+) ghosts (\d+) F0 (\d+)", r.stdout)
    assert m and (int(m.group(1)), int(m.group(2)), int(m.group(3))) == (g.local_vtx_cnt, g.src_ghost_cnt, 9), r.stdout
    for p in range(3):
        assert open(d + "graph.%d.bin" % p, "rb").read() == ds.images[p]
    cache = np.fromfile(d + "feats9.%d.bin" % me, dtype=np.float32).reshape(-1, 9)
    assert np.array_equal(cache, np.concatenate([ds.feats[g.local_to_global], ds.feats[g.src_ghost_gvid]]))

    def checksum(slots):
        return int(sum((int(s) + 1) * (i + 1) for i, s in enumerate(slots)))

    plans = {(int(a), int(b)): tuple(int(x) for x in rest) for a, b, *rest in
             re.findall(r"plan: dir (\d) peer (\d+) recv (\d+) (\d+) send (\d+) (\d+)", r.stdout)}
    assert len(plans) == 4
    for q in (0, 2):
        gq = ds.graphs[q]
        for dname, (mine_ghosts, their_ghosts, my_send, their_send) in enumerate((
                (g.src_ghost_gvid, gq.src_ghost_gvid, g.fwd_send, gq.fwd_send),
                (g.dst_ghost_gvid, gq.dst_ghost_gvid, g.bwd_send, gq.bwd_send))):
            recv = recv_slots(mine_ghosts, gq.local_to_global[their_send[me]])
            send = recv_slots(their_ghosts, g.local_to_global[my_send[q]])
            assert plans[(dname, q)] == (recv.size, checksum(recv), send.size, checksum(send)), (dname, q)


@pytest.mark.gpu
@pytest.mark.parametrize("lambdas", [1, 3])
def test_pipeline_mode_matches_oracle(oracle, lambdas):
    """--pipeline 1: the reference's chunk queues drive the engine (several chunks per partition
    aggregate their own destination ranges); accuracy / loss per epoch as the weight server logs
    them, the epoch and <EM> lines of the graph server, early stop at the target accuracy."""
    from oracle.driver import OracleGCN

    ds = random_dataset(V=900, E_und=7000, dims=[50, 16, 6], seed=78)
    cmd = write_dataset(ds)
    r = subprocess.run(cmd + ["--numepochs", "5", "--pipeline", "1", "--numlambdas", str(lambdas)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [(int(m.group(1)), float(m.group(2)), float(m.group(3)))
           for m in re.finditer(r"Epoch (\d+), acc: ([0-9.]+), loss: ([0-9.]+)", r.stderr)]
    assert [g[0] for g in got] == [1, 2, 3, 4, 5]
    orc = OracleGCN(oracle, ds.graphs, ds.dims)
    orc.load_features(ds.feats, ds.onehot)
    val = int(ds.V * 0.1)
    accs = []
    for ep, acc, loss in got:
        w = orc.epoch()
        accs.append(w["acc"][0] / val)
        assert abs(acc - w["acc"][0] / val) < 2e-4 and abs(loss - w["loss"][0] / val) < 2e-4  # 4 decimals
    assert len(re.findall(r"Sync Epoch \d+ starts", r.stderr)) == 5
    assert len(re.findall(r"Time for epoch \d+: [0-9.]+ms", r.stderr)) == 5
    assert "<EM>: Using %d lambdas" % lambdas in r.stderr and "<EM>: Average  sync epoch time" in r.stderr
    assert "Final: epochs 5, state EARLY" in r.stdout
    # early stop: a target the third epoch reaches -> DONE, no fourth epoch
    target = accs[2]
    if target > max(accs[:2]) and target > 0:
        r = subprocess.run(cmd + ["--numepochs", "5", "--pipeline", "1", "--targetacc", "%.6f" % (target - 1e-4),
                                  "--switchthreshold", "0.0"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "Final: epochs 3, state DONE" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_driver_epochs_match_oracle(oracle):
    from oracle.driver import OracleGCN

    ds = random_dataset(V=900, E_und=7000, dims=[50, 16, 6], seed=77)
    root = tempfile.mkdtemp()
    d = os.path.join(root, "parts_1") + "/"
    os.makedirs(d)
    formats.write_bsnap_edges(d + "graph.bsnap.edges", ds.V, ds.src, ds.dst)
    formats.write_parts(d + "graph.bsnap.parts", np.zeros(ds.V, np.int32))
    formats.write_features(os.path.join(root, "features.bsnap"), ds.feats)
    formats.write_labels(os.path.join(root, "labels.bsnap"), ds.labels, ds.dims[-1])
    formats.write_layer_config(os.path.join(root, "layers.config"), ds.dims)
    r = subprocess.run([BIN, "--datasetdir", d, "--featuresfile", os.path.join(root, "features.bsnap"),
                        "--labelsfile", os.path.join(root, "labels.bsnap"), "--layerfile",
                        os.path.join(root, "layers.config"), "--numepochs", "4"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = [(int(m.group(1)), float(m.group(2)), float(m.group(3)))
           for m in re.finditer(r"Epoch (\d+), acc: ([0-9.]+), loss: ([0-9.]+)", r.stdout)]
    assert len(got) == 4 and os.path.exists(d + "graph.0.bin")
    orc = OracleGCN(oracle, ds.graphs, ds.dims)
    orc.load_features(ds.feats, ds.onehot)
    val = int(ds.V * 0.1)
    for ep, acc, loss in got:
        w = orc.epoch()
        assert abs(acc - w["acc"][0] / val) < 2e-3  # printed with 3 decimals
        assert abs(loss - w["loss"][0] / val) < 2e-3
    # the preprocessed partition the driver wrote is the reference's file, byte for byte
    assert open(d + "graph.0.bin", "rb").read() == ds.images[0]
