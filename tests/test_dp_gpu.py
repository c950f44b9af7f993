"""Multi-GPU gradient equivalence (needs >= 2 GPUs; skipped on a 1-GPU box): see dp_equivalence_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_data_parallel_gradients_equal_microbatch_sum(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dp_equivalence_worker.py")]
    env = dict(os.environ, TRIS_DP_OVERLAP="1")       # exercise the overlapped (two-call) exchange in eager mode
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "DP grads == sum of micro-batch grads: True" in r.stdout
