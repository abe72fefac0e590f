"""spectraldns_b200: B200-native hot path of spectralDNS."""
import os

# The multi-GPU pipeline runs copy streams beside the plan stream, and PyTorch / NCCL bring their own: with the
# default of 8 hardware queues, streams alias and a copy stream waiting for its event stalls kernels queued behind it.
# Must be set before the CUDA context exists.
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
