"""Multi-GPU (slab decomposition over NVLink peer memory) parity: needs >= 2 GPUs on the box;
skipped otherwise.  Launches tests/mp/slab_worker.py under torchrun, one process per GPU."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize('world', [2, 4, 8])
def test_slab_parity(world):
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29500 + world),
           os.path.join(ROOT, 'tests', 'mp', 'slab_worker.py')]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and 'SLAB_WORKER_RESULT fails=0' in r.stdout, r.stdout[-6000:]
