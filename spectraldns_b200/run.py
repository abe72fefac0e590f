"""Run a spectralDNS demo/test script unchanged on the B200 path:

    python -m spectraldns_b200.run /path/to/demo/TG.py [script arguments ...]

Puts spectraldns_b200/compat first on sys.path so that `spectralDNS`, `shenfun`, `mpi4py`,
`mpi4py_fft` (and `h5py` when it is not installed) resolve to the B200 implementations."""
import os
import runpy
import sys

COMPAT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'compat')


def init_distributed():
    """Under torchrun (one process per GPU): bind the GPU and create the NCCL process group that
    stands in for MPI.COMM_WORLD (reductions, rendezvous of the CUDA-IPC handles)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', os.environ.get('RANK', '0')))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))


def activate():
    init_distributed()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (COMPAT, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    for mod in ('spectralDNS', 'shenfun', 'mpi4py', 'mpi4py_fft'):
        if mod in sys.modules and COMPAT not in (getattr(sys.modules[mod], '__file__', '') or ''):
            raise RuntimeError('%s was already imported from elsewhere' % mod)


if __name__ == '__main__':
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    activate()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name='__main__')
