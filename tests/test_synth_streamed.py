"""Per-rank (streamed) generation + dory_preprocess_incident_edges: a rank's edge list is the whole
list filtered by incidence, and the partition image built from it (with the all-gathered in-degrees)
is byte-identical to the one DataLoader::preprocess's restatement builds from the whole list."""
import numpy as np
import pytest

from dorylus_b200 import engine as dengine
from dorylus_b200 import formats, synth

SPEC = synth.GraphSpec("streamed-test", 20_011, 20_011 * 24, [16, 48, 51], seed=5, sigma=0.9, locality=0.9, communities=37)


@pytest.fixture(scope="module")
def whole():
    src, dst, deg, (lo, hi) = synth.generate_incident_edges(SPEC, 0, 1)
    assert (lo, hi) == (0, SPEC.num_vertices)
    return src, dst, deg


def test_whole_graph_shape(whole, monkeypatch):
    src, dst, deg = whole
    assert src.size == SPEC.num_edges and not (src == dst).any()
    assert np.array_equal(src[0::2], dst[1::2]) and np.array_equal(dst[0::2], src[1::2])  # both directions
    assert np.array_equal(deg, np.bincount(dst, minlength=SPEC.num_vertices))
    blk = synth.StreamedLayout(SPEC).blk
    inside = (src // blk) == (dst // blk)
    assert 0.88 < inside.mean() < 0.93  # locality 0.9 (+ remote edges that fall into the same block)
    assert deg.max() > 8 * deg.mean()   # log-normal tail


@pytest.mark.parametrize("world", [2, 5])
def test_rank_lists_are_the_whole_list_filtered(whole, world, monkeypatch):
    monkeypatch.setattr(synth, "_CHUNK_EDGES", 20_000)  # several chunks per rank
    src, dst, deg, _ = synth.generate_incident_edges(SPEC, 0, 1)
    parts = synth.contiguous_parts(SPEC.num_vertices, world)
    degs = []
    for r in range(world):
        s, d, dg, (lo, hi) = synth.generate_incident_edges(SPEC, r, world, threads=2 if r else 1)
        keep = (parts[src] == r) | (parts[dst] == r)
        assert np.array_equal(s, src[keep]) and np.array_equal(d, dst[keep])
        assert np.array_equal(dg, deg[lo:hi])
        degs.append(dg)
        image = dengine.preprocess_incident_edges(s, d, parts, SPEC.num_vertices, r, world, deg, src.size)
        want = dengine.preprocess_edges(src, dst, parts, SPEC.num_vertices, r, world)
        assert bytes(image) == bytes(want)
    assert np.array_equal(np.concatenate(degs), deg)  # what the ranks all-gather


def test_incident_preprocess_rejects_wrong_degrees(whole):
    src, dst, deg = whole
    parts = synth.contiguous_parts(SPEC.num_vertices, 2)
    bad = deg.copy()
    bad[3] += 1
    with pytest.raises(dengine.DoryError):
        dengine.preprocess_incident_edges(src, dst, parts, SPEC.num_vertices, 0, 2, bad, src.size)


def test_feature_and_label_rows_do_not_depend_on_the_range():
    full = synth.generate_feature_rows(0, 5000, 7, seed=3, rows_per_chunk=1024)
    assert np.array_equal(synth.generate_feature_rows(1500, 4100, 7, seed=3, rows_per_chunk=1024), full[1500:4100])
    lab = synth.generate_label_rows(0, 5000, 51, seed=4, rows_per_chunk=1024)
    assert np.array_equal(synth.generate_label_rows(1023, 2049, 51, seed=4, rows_per_chunk=1024), lab[1023:2049])
    assert lab.max() < 51 and np.abs(full).max() <= 1.0
