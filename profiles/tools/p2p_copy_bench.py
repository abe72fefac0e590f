"""Copy-engine peer copy rates on this box: 1-D vs strided 2-D (the slab-exchange shapes), 1 vs 2 streams.
Run on >= 2 GPUs: python profiles/tools/p2p_copy_bench.py"""
import ctypes as C
import torch

rt = C.CDLL('libcudart.so.12')
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
D2D = 3
a = torch.empty(1 << 29, dtype=torch.uint8, device='cuda:0')
b = torch.empty(1 << 29, dtype=torch.uint8, device='cuda:1')
b.copy_(a)      # enables peer access
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
torch.cuda.set_device(0)
streams = [torch.cuda.Stream(device=0) for _ in range(4)]


def timed(fn, nbytes, reps=5):
    best = 0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(0)
        e0.record(torch.cuda.current_stream())
        for s in streams:
            s.wait_event(e0)
        fn()
        for s in streams:
            ev = torch.cuda.Event(); ev.record(s); torch.cuda.current_stream().wait_event(ev)
        e1.record(torch.cuda.current_stream())
        torch.cuda.synchronize(0)
        best = max(best, nbytes/e0.elapsed_time(e1)*1e-6)
    return best


def c1d(n, ns=1):
    def f():
        for i in range(ns):
            o = i*(n//ns)
            rt.cudaMemcpyAsync(b.data_ptr() + o, a.data_ptr() + o, n//ns, D2D, streams[i].cuda_stream)
    return timed(f, n)


def c2d(width, height, spitch, dpitch, ns=1):
    def f():
        h = height//ns
        for i in range(ns):
            rt.cudaMemcpy2DAsync(b.data_ptr() + i*h*dpitch, dpitch, a.data_ptr() + i*h*spitch, spitch, width, h, D2D,
                                 streams[i].cuda_stream)
    return timed(f, width*height)


for n in (8 << 20, 32 << 20, 180 << 20):
    print('1-D %4d MiB: 1 stream %.0f GB/s, 2 streams %.0f GB/s, 4 streams %.0f GB/s' % (n >> 20, c1d(n), c1d(n, 2), c1d(n, 4)))
for width, height in ((352, 65536), (29*1024, 1536), (117*1024, 1536), (33*1024, 768), (132*1024, 768)):
    sp, dp = width, 2*width
    print('2-D rows %6d B x %5d (src dense, dst pitch 2x): 1 stream %.0f GB/s, 2 streams %.0f GB/s, 4 streams %.0f GB/s'
          % (width, height, c2d(width, height, sp, dp), c2d(width, height, sp, dp, 2), c2d(width, height, sp, dp, 4)))
