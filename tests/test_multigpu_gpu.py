"""Runs tests/mgpu_check.py under torchrun on 2 GPUs when the box has them (NCCL path)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_samplers():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29611',
           os.path.join(ROOT, 'tests', 'mgpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'MGPU_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
