"""Row-block sharded V-cycle on >= 2 GPUs of one node: parity with the single-GPU path
(bit-identical for damped Jacobi) and with the oracle.  Skipped on a 1-GPU box."""
import importlib
import os
import subprocess
import sys

import pytest

amg = importlib.import_module("algebraic-multigrid_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("halo", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_vcycle_matches_single_gpu(world, halo):
    if amg.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29511 + world + (10 if halo == "nccl" else 0)),
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    env = dict(os.environ, AMGB_HALO=halo)
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED PARITY OK" in out.stdout
