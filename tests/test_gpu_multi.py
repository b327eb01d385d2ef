"""Multi-GPU parity inside `pytest -m gpu`: launches tests/multi_gpu_check.py (CUDA operator + NCCL ghost
exchange on the Morton box partition against the CPU oracle on the union of the parts) under torchrun on as
many GPUs as the box has; skipped on a 1-GPU box.  The logs of the hand-launched N = 2 / 4 / 8 runs are kept
under profiles/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_parity(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    env = dict(os.environ, GLSB_CHECK_CELLS="32")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" OK") >= 3 * world
