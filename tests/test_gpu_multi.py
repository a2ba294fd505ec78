"""-m gpu, needs >= 2 GPUs (skipped otherwise): the fused gradient exchange + Adam kernel across real ranks, and the whole
data-parallel training step through it against the ncclAllReduce + Adam path (tools/peer_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_peer_exchange_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "tools", "peer_check.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300, cwd=ROOT).stdout.decode()
    assert out.count("PEER_CHECK_OK") == 2, out[-3000:]
    assert "bit-identical across ranks: False" not in out and "timed_out: True" not in out
