"""Partition preprocessor / graph.<id>.bin parser: byte parity with the reference loader."""
import os
import tempfile

import numpy as np
import pytest

from dorylus_b200 import engine as dengine
from dorylus_b200 import formats


def test_images_match_reference_golden(golden):
    """Fixtures were written by the reference's own DataLoader::preprocess (tests/golden/make_golden.py)."""
    r = golden["reference_runs"]
    V, P = int(r["ld_V"]), int(r["ld_P"])
    for und in (0, 1):
        for p in range(P):
            want = r["graph_u%d_p%d" % (und, p)].tobytes()
            got = dengine.preprocess_edges(r["ld_src"], r["ld_dst"], r["ld_parts"], V, p, P, bool(und))
            assert got == want, (und, p)
    want = r["single_graph"].tobytes()
    got = dengine.preprocess_edges(r["single_src"], r["single_dst"], np.zeros(40, np.int32), 40, 0, 1, False)
    assert got == want


def test_parse_roundtrip_and_invariants(golden):
    r = golden["reference_runs"]
    g = formats.parse_graph_bin(r["graph_u0_p1"].tobytes())
    V = g.local_vtx_cnt
    assert g.col_ptrs[0] == 0 and g.col_ptrs[-1] == g.local_in_edge_cnt == g.row_idxs.size
    assert g.row_ptrs[-1] == g.local_out_edge_cnt == g.col_idxs.size
    assert (np.diff(g.src_ghost_gvid.astype(np.int64)) > 0).all()  # ghost slots ascend by global id
    assert np.array_equal(g.src_ghost_lvid, V + np.arange(g.src_ghost_cnt))
    assert np.array_equal(g.dst_ghost_lvid, V + np.arange(g.dst_ghost_cnt))
    assert g.row_idxs.max(initial=0) < V + g.src_ghost_cnt
    deg = np.diff(g.col_ptrs).astype(np.float64) + 1
    assert np.allclose(g.norms, 1.0 / deg, rtol=1e-6)
    assert g.fwd_send[1].size == 0 and g.bwd_send[1].size == 0  # nothing sent to self


def test_live_against_compiled_reference(ref):
    rng = np.random.default_rng(77)
    V, P = 500, 4
    src = rng.integers(0, V, 6000).astype(np.uint32)
    dst = rng.integers(0, V, 6000).astype(np.uint32)
    parts = rng.integers(0, P, V).astype(np.int32)
    d = tempfile.mkdtemp() + "/"
    formats.write_bsnap_edges(d + "graph.bsnap.edges", V, src, dst)
    formats.write_parts(d + "graph.bsnap.parts", parts)
    for p in range(P):
        f = ref.preprocess(d, p, P, False)
        want = open(f, "rb").read()
        assert dengine.preprocess_edges(src, dst, parts, V, p, P, False) == want
        os.remove(f)
        assert open(dengine.preprocess_dir(d, p, P, False), "rb").read() == want
        r = ref.load_graph(d + "graph.%d.bin" % p)  # the reference's Graph::init reads our file
        g = formats.parse_graph_bin(want)
        for k in ("col_ptrs", "row_idxs", "fwd_vals", "row_ptrs", "col_idxs", "bwd_vals", "norms", "local_to_global"):
            assert np.array_equal(getattr(g, k), r[k]), k


def test_randomized_differential_against_compiled_reference(ref):
    """40 small random inputs with everything the format allows at once -- self loops, duplicate
    records, partitions that own no vertex, empty edge lists, single-vertex graphs, both settings of
    --undirected -- through the reference's own DataLoader::preprocess and through ours: every
    graph.<id>.bin must be the same bytes."""
    import shutil

    cases = 0
    for seed in range(40):
        rng = np.random.default_rng(seed)
        V, E, P = int(rng.integers(1, 60)), int(rng.integers(0, 300)), int(rng.integers(1, 6))
        src = rng.integers(0, V, E).astype(np.uint32)
        dst = rng.integers(0, V, E).astype(np.uint32)
        if E > 10:
            src[:3] = dst[:3]                              # self loops (dropped on read)
            src[3:6], dst[3:6] = src[6:9], dst[6:9]        # duplicates (kept, count toward the degree)
        owners = rng.choice(P, size=int(rng.integers(1, P + 1)), replace=False)  # the others own nothing
        parts = owners[rng.integers(0, owners.size, V)].astype(np.int32)
        d = tempfile.mkdtemp() + "/"
        try:
            formats.write_bsnap_edges(d + "graph.bsnap.edges", V, src, dst)
            formats.write_parts(d + "graph.bsnap.parts", parts)
            for p in range(P):
                f = ref.preprocess(d, p, P, bool(seed & 1))
                want = open(f, "rb").read()
                os.remove(f)
                assert dengine.preprocess_edges(src, dst, parts, V, p, P, bool(seed & 1)) == want, (seed, p)
                cases += 1
        finally:
            shutil.rmtree(d, ignore_errors=True)
    assert cases > 100


def test_edge_cases():
    # empty edge list, a partition without local edges, isolated vertices
    parts = np.array([0, 0, 1, 1], np.int32)
    img = dengine.preprocess_edges(np.zeros(0, np.uint32), np.zeros(0, np.uint32), parts, 4, 0, 2)
    g = formats.parse_graph_bin(img)
    assert g.local_vtx_cnt == 2 and g.local_in_edge_cnt == 0 and (g.norms == 1.0).all()
    # only self loops -> all dropped
    img = dengine.preprocess_edges(np.array([1, 2], np.uint32), np.array([1, 2], np.uint32), parts, 4, 1, 2)
    g = formats.parse_graph_bin(img)
    assert g.global_edge_cnt == 0 and g.src_ghost_cnt == 0
    # a cross edge 0 -> 3 creates one dst ghost on part 0 and one src ghost on part 1
    img0 = dengine.preprocess_edges(np.array([0], np.uint32), np.array([3], np.uint32), parts, 4, 0, 2)
    img1 = dengine.preprocess_edges(np.array([0], np.uint32), np.array([3], np.uint32), parts, 4, 1, 2)
    g0, g1 = formats.parse_graph_bin(img0), formats.parse_graph_bin(img1)
    assert g0.dst_ghost_gvid.tolist() == [3] and g0.fwd_send[1].tolist() == [0] and g0.col_idxs.tolist() == [2]
    assert g1.src_ghost_gvid.tolist() == [0] and g1.bwd_send[0].tolist() == [1] and g1.row_idxs.tolist() == [2]
    # value = (indeg(0)+1)^-1/2 (indeg(3)+1)^-1/2 = 1 * 2^-1/2
    assert abs(float(g1.fwd_vals[0]) - 2 ** -0.5) < 1e-7 and g0.bwd_vals[0] == g1.fwd_vals[0]


def test_bad_inputs_raise():
    with pytest.raises(dengine.DoryError):
        dengine.preprocess_edges(np.array([9], np.uint32), np.array([0], np.uint32), np.zeros(4, np.int32), 4, 0, 1)
    with pytest.raises(dengine.DoryError):
        dengine.preprocess_edges(np.array([0], np.uint32), np.array([1], np.uint32), np.array([0, 5, 0, 0], np.int32), 4, 0, 2)
    with pytest.raises(ValueError):
        formats.parse_graph_bin(b"\x00" * 10)


def test_read_features_and_labels_like_the_reference(tmp_path):
    """dory_read_features / dory_read_labels == Engine::readFeaturesFile / readLabelsFile
    (engine/utils.cpp:486-596): rows of local vertices and of source ghosts picked out of the global
    files, the feats<F0>.<id>.bin cache written in the reference's layout (local rows, then ghost
    rows) and -- like the reference -- preferred over the features file once it exists."""
    from helpers import random_dataset
    from dorylus_b200.engine import DoryError, read_features, read_labels

    ds = random_dataset(V=700, E_und=4000, dims=[13, 8, 5], P=3, seed=12)
    d = str(tmp_path) + "/"
    ffile, lfile = d + "features.bsnap", d + "labels.bsnap"
    formats.write_features(ffile, ds.feats)
    formats.write_labels(lfile, ds.labels, 5)
    for p in range(3):
        g = ds.graphs[p]
        loc, gh = read_features(d, ffile, ds.images[p], p, 13)
        assert np.array_equal(loc, ds.feats[g.local_to_global])
        assert np.array_equal(gh, ds.feats[g.src_ghost_gvid])
        cache = d + "feats13.%d.bin" % p
        raw = np.fromfile(cache, dtype=np.float32)
        assert raw.size == (g.local_vtx_cnt + g.src_ghost_cnt) * 13
        assert np.array_equal(raw, np.concatenate([loc.ravel(), gh.ravel()]))
        assert np.array_equal(read_labels(lfile, ds.images[p], 5), ds.onehot[g.local_to_global])
    # the cache wins: change it, the features file is no longer consulted (utils.cpp:488-501)
    g = ds.graphs[1]
    fake = np.arange((g.local_vtx_cnt + g.src_ghost_cnt) * 13, dtype=np.float32)
    fake.tofile(d + "feats13.1.bin")
    loc, gh = read_features(d, d + "does-not-exist", ds.images[1], 1, 13)
    assert np.array_equal(np.concatenate([loc.ravel(), gh.ravel()]), fake)
    # error paths return codes, not asserts
    with pytest.raises(DoryError):
        read_features(d, ffile, ds.images[0], 0, 12)  # width differs from the file's header, no cache for 12
    with pytest.raises(DoryError):
        read_labels(lfile, ds.images[0], 4)  # labelKinds mismatch
    short = d + "short.bsnap"
    formats.write_features(short, ds.feats[:-1])
    with pytest.raises(DoryError):
        read_features(d + "nocache/", short, ds.images[0], 0, 13)
    bad = ds.labels.copy()
    bad[ds.graphs[0].local_to_global[0]] = 9
    formats.write_labels(d + "bad.bsnap", bad, 5)
    with pytest.raises(DoryError):
        read_labels(d + "bad.bsnap", ds.images[0], 5)


def test_numpy_single_partition_loader_equals_the_preprocessor():
    """oracle/np_loader.py (what `bench.py --impl reference` builds its workload with, so that the
    reference arm never loads the product library) against dory_preprocess_edges -- which the tests
    above hold byte-identical to the compiled reference loader: every array of the partition."""
    import dataclasses

    from dorylus_b200 import synth
    from oracle.np_loader import single_partition_graph

    for name in ("reddit-tiny", "cora", "amazon-tiny"):
        spec = synth.CONFIGS[name]
        src, dst = synth.generate_edges(spec)
        src = np.concatenate([src, [1, 2, 3, 3]]).astype(np.uint32)  # self loops (dropped) and a duplicate edge (kept)
        dst = np.concatenate([dst, [1, 2, 4, 4]]).astype(np.uint32)
        V = spec.num_vertices
        got = single_partition_graph(src, dst, V)
        want = formats.parse_graph_bin(dengine.preprocess_edges(src, dst, np.zeros(V, np.int32), V, 0, 1))
        for f in dataclasses.fields(got):
            a, b = getattr(got, f.name), getattr(want, f.name)
            if isinstance(a, np.ndarray):
                assert a.dtype == b.dtype and np.array_equal(a, b), (name, f.name)
            elif isinstance(a, list):
                assert all(np.array_equal(x, y) for x, y in zip(a, b)), (name, f.name)
            else:
                assert a == b, (name, f.name)
