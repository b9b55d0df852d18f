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
