"""Multi-GPU data paths on real GPUs (skipped on a single-GPU box): the NCCL scatter/solve/gather pipeline, the fused
solve + gather over NVLink peer memory and the copy-engine pull / solve / push pipeline all reproduce a single-GPU
solve bit for bit (tools/peer_gather_check.py)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_scatter_gather_paths_bitwise():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tools", "peer_gather_check.py")], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])
    assert "peer_gather_check ok" in out.stdout
