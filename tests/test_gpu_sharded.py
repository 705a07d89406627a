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
@pytest.mark.parametrize("halo", ["peer", "peer_unfused", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_vcycle_matches_single_gpu(world, halo):
    """peer: the legs push their boundary rows into the neighbours' ghost rows themselves;
    peer_unfused: stand-alone peer-memory exchange kernels; nccl: grouped send/recv."""
    if amg.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29511 + world + {"peer": 0, "peer_unfused": 20, "nccl": 10}[halo]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    env = dict(os.environ, AMGB_HALO="nccl" if halo == "nccl" else "peer",
               AMGB_FUSED_PUSH="0" if halo == "peer_unfused" else "1")
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "SHARDED PARITY OK" in out.stdout
