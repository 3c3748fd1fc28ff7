"""NCCL, one process per GPU: scatter -> fit -> gather on real devices.  Needs >= 2 GPUs (skipped on
a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu` runs it)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize('total', [70, 3])
def test_scatter_fit_gather_nccl(total):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if total > 3 else min(n, 4)  # total < world: empty shards
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'dist_gpu_worker.py'), str(total)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('DIST_RESULT ')]
    assert lines, r.stdout[-2000:]
    res = json.loads(lines[-1][len('DIST_RESULT '):])
    assert res['ok'], res
