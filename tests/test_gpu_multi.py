"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise).  Spawns one rank per GPU with
torch.distributed.run and lets tools/multi_gpu_check.py compare every rank with the CPU oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("parts", ["random", "contiguous"])
def test_two_gpu_epochs_match_oracle(parts):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517" if parts == "random" else "29518",
           os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--parts", parts]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout
