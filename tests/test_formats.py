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
    assert (nv, ne) == (4, 3)  # the header counts lines, the body holds both directions
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
