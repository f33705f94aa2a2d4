"""The element-partitioned path on two GPUs of one box (one process per GPU, torchrun): curves and Newton
iterations of the 2-rank analyses against the oracle's single-domain run, shared nodes bit-identical, the
peer-memory halo against the NCCL interface sum.  The work is done by scripts/mgpu_check.py; here it is launched the
way the driver launches bench.py and its verdict (exit code) is the test."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(world, env_extra=None):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "mgpu_check.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_rank_analysis_matches_single_domain(p2p):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    p = _launch(2, {"FCVM_P2P": p2p})
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    assert "FAIL" not in p.stdout
    if p2p == "1":
        assert "bit-identical=True" in p.stdout
