"""Text -> binary converters of the reference's inputs/ directory (formats.convert_*_text) and the
round trip through the loader.  CPU only."""
import numpy as np

from dorylus_b200 import formats
from dorylus_b200.engine import preprocess_dir


def test_graph_text_to_bsnap(tmp_path):
    txt = tmp_path / "g.txt"
    txt.write_text("# comment\n% another\n0 1\n1 2\n2 2\n3 0\n4\tx\n5 6\n")  # self loop dropped, stops at '4 x'
    out = str(tmp_path / "graph.bsnap")
    nv, ne = formats.convert_graph_text(str(txt), out, undirected=False)
    assert (nv, ne) == (4, 3)
    v, s, d = formats.read_bsnap_edges(out)
    assert v == 4 and s.tolist() == [0, 1, 3] and d.tolist() == [1, 2, 0]
    nv, ne = formats.convert_graph_text(str(txt), out, undirected=True)
    v, s, d = formats.read_bsnap_edges(out)
    assert (nv, ne) == (4, 6)  # header and body both count the two directions (graphToBinary.cpp:151)
    assert s.tolist() == [0, 1, 1, 2, 3, 0] and d.tolist() == [1, 0, 2, 1, 0, 3]
    # the loader reads it
    import os
    d2 = str(tmp_path / "parts_1") + "/"
    os.makedirs(d2)
    os.replace(out, d2 + "graph.bsnap.edges")
    formats.write_parts(d2 + "graph.bsnap.parts", np.zeros(4, np.int32))
    g = formats.read_graph_bin(preprocess_dir(d2, 0, 1))
    assert g.local_vtx_cnt == 4 and g.local_in_edge_cnt == 6


def test_features_and_labels_text_to_bsnap(tmp_path):
    f = tmp_path / "features"
    f.write_text("0.5, 1.25,2\n  3 4 5  \n-1, 2, 3\n\nnot a row\n7,8,9")
    rows, skipped = formats.convert_features_text(str(f), 3)
    assert (rows, skipped) == (3, 2)  # the row that starts with '-' is dropped, like the reference does
    got = formats.read_features(str(f) + ".bsnap")
    assert np.array_equal(got, np.array([[0.5, 1.25, 2], [3, 4, 5], [7, 8, 9]], np.float32))
    l = tmp_path / "labels"
    l.write_text("3\n0\n\n x\n12abc\n1")
    n, skipped = formats.convert_labels_text(str(l), 13)
    kinds, lab = formats.read_labels(str(l) + ".bsnap")
    assert (n, skipped, kinds) == (4, 1, 13) and lab.tolist() == [3, 0, 12, 1]


GRAPH_TEXTS = {
    "plain": "0 1\n1 2\n2 0\n",
    "comments-loops-stop": "# comment\n% another\n0 1\n1 2\n2 2\n3 0\n4\tx\n5 6\n",
    "tabs-and-gaps": "10\t3\n3 10\n\n7 7\n0 99\n",
    "empty": "# nothing\n",
}


def _ref_tools():
    from oracle import build as ob

    return ob.build_ref_tools()


def test_graph_text_converter_matches_the_reference_tool(tmp_path):
    """inputs/graphToBinary.cpp compiled as it is (oracle/_ref/graphToBinary) on the same text files:
    our converter writes the same bytes, header included, for --undirected 0 and 1."""
    import subprocess

    tools = _ref_tools()
    if "graphToBinary" not in tools:
        import pytest

        pytest.skip("oracle/_ref/graphToBinary not available")
    for name, text in GRAPH_TEXTS.items():
        for und in (0, 1):
            t = tmp_path / ("%s_%d.txt" % (name, und))
            t.write_text(text)
            r = subprocess.run([tools["graphToBinary"], "--snapfile=%s" % t, "--undirected=%d" % und, "--header=1"],
                               capture_output=True, text=True, timeout=60)
            assert r.returncode == 0, r.stderr
            want = open(str(t) + ".bsnap", "rb").read()
            ours = str(tmp_path / "ours.bsnap")
            formats.convert_graph_text(str(t), ours, undirected=bool(und))
            assert open(ours, "rb").read() == want, (name, und)


def test_readers_parse_what_the_reference_generators_write(tmp_path):
    """inputs/generateFeatues.cpp and generateLabels.cpp (compiled as they are) write features / labels
    files directly in the binary layout the graph server reads: our readers (Python and the C++
    dory_read_features / dory_read_labels) must see exactly those values."""
    import subprocess

    import pytest

    from dorylus_b200 import engine as dengine

    tools = _ref_tools()
    if "generateFeatues" not in tools or "generateLabels" not in tools:
        pytest.skip("oracle/_ref generators not available")
    V, F, K = 37, 24, 5
    base = str(tmp_path / "ds")
    for tool, args in (("generateFeatues", [str(V), str(F), base]), ("generateLabels", [str(V), str(K - 1), base])):
        r = subprocess.run([tools[tool]] + args, capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
    raw = np.fromfile(base + ".feats", dtype=np.uint8)
    assert int(raw[:4].view(np.uint32)[0]) == F
    want = raw[4:].view(np.float32).reshape(V, F)
    nnz = (want != 0).sum(1)
    assert (nnz >= F // 3).all() and (nnz <= F * 3 // 4).all() and np.abs(want).max() < 1  # generateFeatues.cpp:32-55
    assert np.array_equal(formats.read_features(base + ".feats"), want)
    lraw = np.fromfile(base + ".labels", dtype=np.uint32)
    # generateLabels draws from [0, numLabels] INCLUSIVE (uniform_int_distribution(0, numLabels)): with
    # numLabels = K - 1 every label is a valid class of a K-class model
    assert lraw[0] == K - 1 and lraw[1:].max() <= K - 1
    kinds, lab = formats.read_labels(base + ".labels")
    assert kinds == K - 1 and np.array_equal(lab, lraw[1:])
    # the C++ readers, through a single-partition image
    src = np.arange(V, dtype=np.uint32)
    image = dengine.preprocess_edges(src, (src + 1) % V, np.zeros(V, np.int32), V, 0, 1)
    local, ghost = dengine.read_features(str(tmp_path), base + ".feats", image, 0, F)
    assert np.array_equal(local, want) and ghost.shape[0] == 0
    # readLabelsFile asserts header == layer config and label < labelKinds (engine/utils.cpp:567,582), so
    # the generator's own file is only readable when the inclusive upper label was never drawn
    assert lraw[1:].max() == K - 1  # it is drawn in this (deterministic) sample
    with pytest.raises(dengine.DoryError):
        dengine.read_labels(base + ".labels", image, K)       # header says K - 1
    with pytest.raises(dengine.DoryError):
        dengine.read_labels(base + ".labels", image, K - 1)   # a label equals labelKinds: the reference asserts
    fixed = lraw.copy()
    fixed[0] = K  # what a user has to do: declare one class more than the generator was asked for
    fixed.tofile(base + ".labels")
    onehot = dengine.read_labels(base + ".labels", image, K)
    assert np.array_equal(onehot.argmax(1), lraw[1:]) and (onehot.sum(1) == 1).all()


def test_converter_and_readers_against_golden_tool_outputs(golden, tmp_path):
    """The same two comparisons from the committed outputs of the reference tools
    (tests/golden/input_tools.npz), for machines without oracle/_ref."""
    g = golden["input_tools"]
    for name, text in GRAPH_TEXTS.items():
        for und in (0, 1):
            t = tmp_path / "g.txt"
            t.write_text(text)
            ours = str(tmp_path / "ours.bsnap")
            formats.convert_graph_text(str(t), ours, undirected=bool(und))
            assert open(ours, "rb").read() == g["bsnap_%s_%d" % (name, und)].tobytes(), (name, und)
    f = tmp_path / "ds.feats"
    f.write_bytes(g["gen_feats"].tobytes())
    want = np.frombuffer(g["gen_feats"].tobytes()[4:], dtype=np.float32).reshape(37, 24)
    assert np.array_equal(formats.read_features(str(f)), want)
    l = tmp_path / "ds.labels"
    l.write_bytes(g["gen_labels"].tobytes())
    kinds, lab = formats.read_labels(str(l))
    raw = np.frombuffer(g["gen_labels"].tobytes(), dtype=np.uint32)
    assert kinds == 4 and np.array_equal(lab, raw[1:])
