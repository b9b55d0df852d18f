"""The apply-first schedule (A_hat . (in . W) where a layer narrows) reproduces the reference epoch:
checked on the CPU between the oracle's reference epoch and tests/apply_first_model.py, for one and
several partitions and for mixed per-layer choices (the GPU path is checked against the reference
oracle in tests/test_gpu_apply_first.py)."""
import numpy as np
import pytest

from apply_first_model import ApplyFirstGCN, choose_apply_first
from helpers import random_dataset, rel_err
from oracle.driver import OracleGCN


def test_rule_picks_narrowing_layers():
    assert choose_apply_first([602, 128, 41]) == [True, True]           # Reddit: 608 -> 128 -> 64
    assert choose_apply_first([100, 64, 64, 25]) == [True, False, True]  # Amazon: 128 -> 64 -> 64 -> 32
    assert choose_apply_first([16, 48, 51]) == [False, False]            # Friendster: 16 -> 64 -> 64
    assert choose_apply_first([1433, 16, 7]) == [True, True]             # Cora


@pytest.mark.parametrize("dims,P,af", [([40, 12, 5], 1, None), ([40, 12, 5], 3, None), ([24, 16, 16, 4], 2, [True, False, True]),
                                       ([24, 16, 16, 4], 2, [False, True, False]), ([12, 20, 6], 2, [False, True]),
                                       ([12, 20, 6], 1, [True, False])])
def test_apply_first_schedule_equals_reference_epoch(oracle, dims, P, af):
    ds = random_dataset(V=260, E_und=1700, dims=dims, P=P, seed=5)
    ref, mod = OracleGCN(oracle, ds.graphs, dims), ApplyFirstGCN(oracle, ds.graphs, dims, apply_first=af)
    ref.load_features(ds.feats, ds.onehot)
    mod.load_features(ds.feats, ds.onehot)
    for ep in range(3):
        a, b = ref.epoch(), mod.epoch()
        assert a["acc"] == b["acc"]
        assert np.allclose(a["loss"], b["loss"], rtol=1e-5)
        L = len(dims) - 1
        for p in range(P):
            for l in range(L - 1):
                assert rel_err(mod.saved[p][l]["z"], ref.saved[p][l]["z"]) < 1e-5, (ep, p, l, "z")
                assert rel_err(mod.saved[p][l]["h"], ref.saved[p][l]["h"]) < 1e-5, (ep, p, l, "h")
                assert rel_err(mod.saved[p][l]["aTg"], ref.saved[p][l]["aTg"]) < 2e-5, (ep, p, l, "aTg")
        for l in range(L):
            tot_r = sum(ref.dW[p][l] for p in range(P))
            tot_m = sum(mod.dW[p][l] for p in range(P))
            assert rel_err(tot_m, tot_r) < 2e-5, (ep, l, "dW")
        for l in range(L):  # same weights into the next epoch
            mod.W[l][:] = ref.W[l]
