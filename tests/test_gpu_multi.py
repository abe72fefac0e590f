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


@pytest.mark.parametrize('layout', ['blocks', 'cyclic'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_slab_parity(world, layout):
    """layout: which axis-1 modes a rank owns -- the reference's contiguous blocks, or [rank::P] (Plan(k1_layout='cyclic'),
    the balanced ownership bench.py uses on 3 or more GPUs under the 2/3 rule)."""
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    env = dict(os.environ, SLAB_K1_LAYOUT=layout)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29500 + world),
           os.path.join(ROOT, 'tests', 'mp', 'slab_worker.py')]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and 'SLAB_WORKER_RESULT fails=0' in r.stdout, r.stdout[-6000:]


def test_reference_tg_script_two_ranks(tmp_path):
    """The reference's tests/TG.py driver, unchanged, on 2 GPUs (the reference's `mpirun -np 2`
    configuration, BASELINE.json configs[0]): its own k / w asserts run on rank 0."""
    if _ngpu() < 2:
        pytest.skip('needs 2 GPUs')
    ref = os.path.join(ROOT, 'baseline', '_ref', 'tests', 'TG.py')
    if not os.path.exists(ref):
        pytest.skip('reference scripts not staged')
    env = dict(os.environ)
    env['PYTHONPATH'] = ROOT + os.pathsep + env.get('PYTHONPATH', '')
    for solver in ('NS', 'VV'):
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
               '--master-addr', '127.0.0.1', '--master-port', '29520', '-m', 'spectraldns_b200.run', ref, solver]
        r = subprocess.run(cmd, cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=600)
        assert r.returncode == 0 and 'Fastest' in r.stdout, r.stdout[:4000] + '\n...\n' + r.stdout[-2000:]
