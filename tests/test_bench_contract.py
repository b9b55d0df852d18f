"""bench.py's reference arm (the CPU oracle timed on the host cores) runs without a GPU: check the
JSON line it prints against the contract the driver reads (one line on stdout, the keys, the
`impl` / `cpu_baseline` / `e2e` shape of the reference arm)."""
import json
import os

import numpy as np
import pytest
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "reddit-small", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "aggregated_edges_per_sec" and d["unit"] == "edges/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # "reference": the reference's own gcn_ops.cpp / CPU_comm.cpp object code (oracle/_ref/librefengine.so) is there
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "reddit-small"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


class _FakeEngine:
    """Stands in for dorylus_b200.engine.Engine so that bench.py's OWN control flow (our arm) can be
    executed without a GPU: every call the bench makes exists here with the same signature and returns
    plausible values; nothing is computed.  This is a test of bench.py's plumbing and of the JSON
    contract, not of the engine."""

    def __init__(self, dims, gnn_type=0, node_id=0, num_nodes=1, device=0, learning_rate=0.01, flags=0):
        self.dims, self.flags, self.nodeId = list(dims), flags, node_id
        self.launches, self.localVtxCnt = 0, 0
        self.calls = []

    def load_partition(self, image):
        from dorylus_b200 import formats

        self.localVtxCnt = formats.parse_graph_bin(image).local_vtx_cnt

    def apply_first(self, layer):
        return bool(self.flags & 0x8) and self.dims[layer + 1] < self.dims[layer]

    def whole_chunk(self, layer=0, dir=0, epoch=1, vertex=True):
        return (layer, dir)

    def _launch(self, name, n=1):
        self.calls.append(name)
        self.launches += n

    def set_tensor(self, layer, name, host):
        self._launch("set_tensor")

    def prefetch_tensor(self, layer, name, host):
        self.calls.append("prefetch")

    def commit_prefetch(self):
        self._launch("commit")

    def scatter(self, chunk):
        self._launch("scatter")

    def aggregate(self, chunk):
        self._launch("aggregate", 2)

    def init_weights(self):
        pass

    def epoch(self):
        self._launch("epoch", 20)
        return self.stats()

    def epoch_async(self):
        self._launch("epoch", 20)

    def stats(self):
        return dict(acc_sum=1.0, loss_sum=2.0, val_rows=1, epochs_done=1, kernel_launches=self.launches, edges_aggregated=1)

    def stats_enqueue(self, slot=0):
        pass

    def stats_collect(self, slot=0):
        return self.stats()

    def sync(self):
        pass

    def event_record(self, slot):
        pass

    def event_elapsed_ms(self, a, b):
        return 1.5

    def measure_fma_peak(self):
        return 70.0

    def close(self):
        pass


def _run_our_arm_with_fake_engine(monkeypatch, capfd, argv, fake_engine=True, child_arm=False):
    import importlib

    import torch

    import dorylus_b200.engine as dengine

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    if fake_engine:
        monkeypatch.setattr(dengine, "Engine", _FakeEngine)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + ([] if child_arm else ["--no-arms"]) + (["--no-invariants"] if fake_engine else []) + argv)
    if child_arm:
        monkeypatch.setenv("DORY_BENCH_ARMS", "reddit_gcn_apply_first")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    saved = os.dup(1)  # bench.py re-points fd 1 at stderr for the libraries it loads
    try:
        assert bench.main() == 0
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    out = capfd.readouterr().out
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_our_arm_control_flow_and_contract_keys(monkeypatch, capfd):
    d = _run_our_arm_with_fake_engine(monkeypatch, capfd, ["--workload", "reddit-tiny", "--steps", "3", "--warmup", "3",
                                                           "--cpu-steps", "1"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["gpu_launches"] == 60 and "impl" not in d
    assert set(d["per_layer_ms"]) == {"L0_fwd", "L1_fwd", "L1_bwd"}
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert "F=602" in r["kernel"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 600 * 602 * 4 + 600 * 41 * 4 and e["d2h_bytes_per_step"] == 8 and e["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["details"]["aggregations_per_step"] == 3 and "reference order" in d["details"]["schedule"]
    assert set(d["config"]) == {"workload", "V", "E", "dims", "parallelism", "l2_policy"}


def test_our_arm_apply_first_flag(monkeypatch, capfd):
    d = _run_our_arm_with_fake_engine(monkeypatch, capfd, ["--workload", "reddit-tiny", "--steps", "2", "--apply-first",
                                                           "--no-cpu-baseline"])
    assert set(d["per_layer_ms"]) == {"L0_fwd", "L1_fwd", "L1_bwd", "L0_bwd"}
    c = d["details"]
    assert c["aggregations_per_step"] == 3 and c["aggregations_launched_per_step"] == 4  # value stays on the reference's job
    assert c["aggregated_row_widths"]["L0_fwd"] == 128 and c["aggregated_row_widths"]["L1_bwd"] == 41
    assert "apply-first on layers [0, 1]" in c["schedule"] and "F=128" in d["roofline"]["kernel"]
    assert d["roofline"]["traffic"] is None and "cpu_baseline" not in d


def test_our_arm_three_layers_on_the_emulated_engine(hostcheck, monkeypatch, capfd):
    """The Amazon widths (3 layers; apply-first picks layers 0 and 2, layer 1 keeps the reference order)."""
    d = _run_our_arm_with_fake_engine(monkeypatch, capfd, ["--workload", "amazon-tiny", "--steps", "2", "--no-cpu-baseline",
                                                           "--apply-first"], fake_engine=False)
    c = d["details"]
    assert c["aggregations_per_step"] == 5 and c["aggregations_launched_per_step"] == 6
    assert set(d["per_layer_ms"]) == {"L0_fwd", "L1_fwd", "L2_fwd", "L2_bwd", "L1_bwd", "L0_bwd"}
    assert c["aggregated_row_widths"]["L0_fwd"] == 64 and c["aggregated_row_widths"]["L1_fwd"] == 64 and c["aggregated_row_widths"]["L2_bwd"] == 25
    assert np.isfinite(d["loss_sum"]) and d["loss_sum"] > 0


@pytest.mark.parametrize("extra", [[], ["--apply-first"]], ids=["reference-order", "apply-first"])
def test_our_arm_on_the_emulated_engine(hostcheck, monkeypatch, capfd, extra):
    """bench.py's own arm against the REAL engine object on the emulated runtime (tests/hostcheck):
    uploads, the input pipeline (prefetch / commit), stream-ordered statistics, per-aggregation timing
    calls -- every ABI call the bench makes is executed, on the tiny Reddit-width workload."""
    d = _run_our_arm_with_fake_engine(monkeypatch, capfd, ["--workload", "reddit-tiny", "--steps", "2", "--no-cpu-baseline"] + extra,
                                      fake_engine=False)
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert np.isfinite(d["loss_sum"]) and d["loss_sum"] > 0
    assert d["details"]["aggregations_launched_per_step"] == (4 if extra else 3)
    inv = d["invariants"]  # checksums of the first epoch, computed for real on the emulated engine
    assert inv["fwd_checksum_rel_err"] < 1e-5 and inv["bwd_checksum_rel_err"] < 1e-5


def test_apply_first_child_arm_cannot_cost_the_headline(monkeypatch, capfd):
    """At N = 1 the bench measures the opt-in apply-first schedule in a CHILD process after the headline
    measurement.  Here the child finds no GPU and fails: the headline line must come out regardless,
    carrying the child's failure under configs.reddit_gcn_apply_first."""
    d = _run_our_arm_with_fake_engine(monkeypatch, capfd, ["--workload", "reddit-tiny", "--steps", "2", "--no-cpu-baseline"],
                                      child_arm=True)
    assert d["value"] > 0 and d["gpu_launches"] == 40 and "reference order" in d["details"]["schedule"]
    arm = d["configs"]["reddit_gcn_apply_first"]
    assert "error" in arm and "exit 2" in arm["error"]
