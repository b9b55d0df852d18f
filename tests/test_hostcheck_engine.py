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
