"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): peer-memory owner update vs NCCL owner update vs
a single-GPU run over the concatenated pair shards.  See tests/multi_gpu_worker.py."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_owner_update_matches_nccl_and_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    g = 2 if n < 4 else 4
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={g}', '--master-addr',
           '127.0.0.1', '--master-port', '29611', os.path.join(ROOT, 'tests', 'multi_gpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'MULTI_GPU_PARITY PASS' in out.stdout
