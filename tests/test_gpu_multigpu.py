"""Multi-GPU correctness of the row-sharded field sum over peer memory, collected by pytest: launches
tests/run_peer_image_multigpu.py under torchrun on 2 GPUs (skipped on a single-GPU box, where the same code runs
as a world of one rank in tests/test_gpu_parity.py::test_peer_image_world_size_one)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_peer_image_two_gpus():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "run_peer_image_multigpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_IMAGE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
