"""Multi-GPU CUDA path (needs >= 2 GPUs; skipped on a single-GPU box): torchrun + NCCL halo exchange, checked
bit for bit against the single-domain oracle by tools/multi_gpu_check.py."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,pgrid", [(2, "1x2"), (2, "2x1"), (4, "2x2")])
def test_multi_gpu_halo_exchange_bit_exact(nproc, pgrid):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--pgrid", pgrid]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BIT-IDENTICAL" in r.stdout
