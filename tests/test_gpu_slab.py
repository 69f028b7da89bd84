"""GPU (>= 2 devices): slab-decomposed TDGL equals the single-GPU run bitwise."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_ngpu() < 2, reason="needs two CUDA devices")
def test_slab_equals_single_gpu_bitwise():
    n = min(_ngpu(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "slab_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SLAB CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
