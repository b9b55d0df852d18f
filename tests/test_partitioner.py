"""The edge-cut partitioner that stands in for inputs/partitioner.cpp (METIS): host only, no GPU.

There is no METIS here to compare against, so the checks are the properties the reference relies
on: every vertex gets exactly one owner in [0, P), the `.parts` / `.comm` files have the format
DataLoader::readPartsFile reads (graph/dataloader.cpp:53-87), partitions are balanced, the result
is deterministic, and on a graph WITH communities the cut is a small fraction of a random
assignment's -- plus the end-to-end property that matters downstream: the images preprocessed from
these owners load and have far fewer ghost rows."""
import os

import numpy as np
import pytest

from dorylus_b200 import formats, synth
from dorylus_b200.engine import DoryError, partition_edges, partition_file, preprocess_edges


def community_graph(V=4000, communities=40, E_und=60000, locality=0.9, seed=5, shuffle=True):
    spec = synth.GraphSpec("c", V, 2 * E_und, [8, 4, 2], seed=seed, sigma=0.7, locality=locality,
                           communities=communities)
    src, dst = synth.generate_edges(spec)
    if shuffle:  # hide the community structure from the vertex numbering
        perm = np.random.default_rng(seed + 1).permutation(V).astype(np.uint32)
        src, dst = perm[src], perm[dst]
    return src, dst


def test_balanced_low_cut_and_deterministic():
    V, P = 4000, 8
    src, dst = community_graph(V=V)
    parts, cut = partition_edges(src, dst, V, P)
    assert parts.dtype == np.int32 and parts.shape == (V,) and parts.min() >= 0 and parts.max() < P
    assert cut == int(np.count_nonzero(parts[src] != parts[dst]))
    sizes = np.bincount(parts, minlength=P)
    assert sizes.max() <= 1.04 * V / P + 1
    indeg = np.bincount(dst, minlength=V)
    edges = np.bincount(parts, weights=indeg, minlength=P)
    assert edges.max() <= 1.12 * src.size / P  # in-edge (= aggregation work) balance
    rnd = synth.random_parts(V, P, seed=1)
    random_cut = int(np.count_nonzero(rnd[src] != rnd[dst]))
    assert cut < 0.35 * random_cut, (cut, random_cut)  # ~10 % of the edges leave their community
    again, cut2 = partition_edges(src, dst, V, P)
    assert np.array_equal(parts, again) and cut == cut2


def test_fewer_ghost_rows_downstream():
    V, P = 3000, 4
    src, dst = community_graph(V=V, communities=24, E_und=40000)
    parts, _ = partition_edges(src, dst, V, P)
    rnd = synth.random_parts(V, P, seed=2)
    ghosts = {}
    for name, owner in (("ours", parts), ("random", rnd)):
        total = 0
        for p in range(P):
            g = formats.parse_graph_bin(preprocess_edges(src, dst, owner, V, p, P))
            assert g.local_vtx_cnt == int(np.count_nonzero(owner == p))
            total += g.src_ghost_cnt
        ghosts[name] = total
    assert ghosts["ours"] < 0.7 * ghosts["random"], ghosts


def test_degenerate_inputs():
    parts, cut = partition_edges(np.zeros(0, np.uint32), np.zeros(0, np.uint32), 10, 3)
    assert parts.min() >= 0 and parts.max() < 3 and cut == 0
    assert np.bincount(parts, minlength=3).max() <= 5
    src, dst = community_graph(V=200, communities=4, E_und=900)
    parts, cut = partition_edges(src, dst, 200, 1)
    assert not parts.any() and cut == 0
    with pytest.raises(DoryError):
        partition_edges(np.array([0, 7], np.uint32), np.array([1, 2], np.uint32), 5, 2)  # vertex 7 >= 5


def test_partition_file_writes_the_reference_formats(tmp_path):
    V, P = 500, 4
    src, dst = community_graph(V=V, communities=8, E_und=3000)
    path = tmp_path / "graph.bsnap"
    formats.write_bsnap_edges(str(path), V, src, dst)
    out = partition_file(str(path), P, str(tmp_path))
    assert out == os.path.join(str(tmp_path), "graph.bsnap.parts")
    lines = open(out).read().split("\n")
    assert lines[-1] == "" and len(lines) == V + 1 and all(l.isdigit() for l in lines[:-1])
    owners = np.array(lines[:-1], dtype=np.int32)
    want, cut = partition_edges(src, dst, V, P)
    assert np.array_equal(owners, want)
    assert open(os.path.join(str(tmp_path), "graph.bsnap.comm")).read() == "Communication cost: %d\n" % cut
    assert np.array_equal(formats.read_parts(out), owners)  # the loader's reader accepts it
