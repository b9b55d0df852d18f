"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise).  Spawns one rank per GPU with
torch.distributed.run and lets tools/multi_gpu_check.py compare every rank with the CPU oracle, for
both ghost-exchange paths (peer-memory store kernel, NCCL all-to-all-v)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("parts,exchange,port", [("random", "p2p", 29517), ("contiguous", "p2p", 29518),
                                                 ("random", "nccl", 29519)])
def test_two_gpu_epochs_match_oracle(parts, exchange, port):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--parts", parts, "--exchange", exchange]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout
