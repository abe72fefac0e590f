// host_shim.h -- TEST INFRASTRUCTURE ONLY.  A small emulation of the CUDA execution model and of the slice of the
// CUDA runtime the library uses, so that g++ can compile spectraldns_b200/csrc/*.cu{,h} for the CPU:
//   * tests/host/fft_core_harness.cpp drives the FFT core with the threads of a line in lockstep;
//   * tests/host/build_emu.py builds the whole library (plan, C ABI, every kernel for a few small transform lengths)
//     into tests/host/build/libsdns_emu.so, in which every CUDA thread of a block is an OS thread, __syncthreads /
//     __syncwarp are barriers, warp shuffles go through a per-warp mailbox, and shared memory is a per-block buffer
//     pre-filled with NaN bytes.  tests/test_kernels_emulated.py runs the parity cases through it on tiny grids.
// It checks kernel LOGIC (indexing, maps, barriers, epilogues) without a GPU.  It is not a backend: the product
// (spectraldns_b200/_lib.py) only ever loads the nvcc-built libsdns_b200.so and fails without a CUDA device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <barrier>
#include <chrono>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __align__(n) alignas(n)
#define __launch_bounds__(...)

struct float2 { float x, y; };
struct double2 { double x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
// the packed f32x2 intrinsics of sm_100: lane-wise, round to nearest
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
template <typename Q> inline Q __ldg(const Q* p) { return *p; }
inline int __float2int_rz(float v) { return (int)std::trunc(v); }
inline long long clock64() { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }   // ns: the flag barrier's 2e10-"cycle" timeout becomes 20 s
inline void __threadfence_system() {}

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };

namespace sdns_emu {
struct Cta {
    std::barrier<> all;
    std::vector<std::unique_ptr<std::barrier<>>> warp;
    std::vector<unsigned long long> mail;            // warp-shuffle mailbox [warp][lane]
    std::vector<unsigned char> dyn, stat;
    Cta(int nthreads, size_t smem) : all(nthreads), mail((size_t)((nthreads + 31) / 32) * 32), dyn(smem + 256, 0xFF), stat(4096, 0xFF) {
        for (int w = 0; w * 32 < nthreads; ++w)
            warp.emplace_back(new std::barrier<>(std::min(32, nthreads - w * 32)));
    }
};
struct Tls { dim3 tid, bid, bdim, gdim; Cta* cta; int lin; };
inline Tls& tls() { static thread_local Tls t; return t; }
inline unsigned char* dyn_smem() { unsigned char* p = tls().cta->dyn.data(); return p + ((256 - (uintptr_t)p % 256) % 256); }
inline unsigned char* static_smem() { unsigned char* p = tls().cta->stat.data(); return p + ((256 - (uintptr_t)p % 256) % 256); }

// per-rank delay before every launch (set by the rank's host thread through sdns_emu_set_skew): makes one rank
// systematically slower than its peers, which turns missing cross-rank ordering into wrong results
inline int& skew_us() { static thread_local int v = 0; return v; }
template <class F> struct Launcher {
    F f; dim3 g, b; size_t smem;
    template <class... A> void operator()(A... args) const {
        const int nt = (int)(b.x * b.y * b.z);
        // SDNS_EMU_JITTER=<max microseconds>: every launch starts after a random delay, which pulls the ranks of an
        // emulated multi-GPU run apart and widens the windows of missing cross-rank ordering
        static const int jitter = std::getenv("SDNS_EMU_JITTER") ? std::atoi(std::getenv("SDNS_EMU_JITTER")) : 0;
        if (skew_us() > 0) std::this_thread::sleep_for(std::chrono::microseconds(skew_us()));
        if (jitter > 0) {
            static thread_local unsigned long long seed = 88172645463325252ULL ^ (unsigned long long)std::hash<std::thread::id>()(std::this_thread::get_id());
            seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
            std::this_thread::sleep_for(std::chrono::microseconds(seed % (unsigned long long)jitter));
        }
        for (unsigned bz = 0; bz < g.z; ++bz) for (unsigned by = 0; by < g.y; ++by) for (unsigned bx = 0; bx < g.x; ++bx) {
            Cta cta(nt, smem);
            std::vector<std::thread> th;
            th.reserve(nt);
            for (int i = 0; i < nt; ++i)
                th.emplace_back([&, i]() {
                    Tls& t = tls();
                    t.tid = dim3(i % b.x, (i / b.x) % b.y, i / (b.x * b.y)); t.bid = dim3(bx, by, bz); t.bdim = b; t.gdim = g;
                    t.cta = &cta; t.lin = i;
                    f(args...);
                    cta.warp[i / 32]->arrive_and_drop();           // exited threads no longer take part in barriers
                    cta.all.arrive_and_drop();
                });
            for (auto& x : th) x.join();
        }
    }
};
template <class F> Launcher<F> launcher(F f, dim3 g, dim3 b, size_t smem) { return Launcher<F>{f, g, b, smem}; }
}  // namespace sdns_emu

#define threadIdx (sdns_emu::tls().tid)
#define blockIdx (sdns_emu::tls().bid)
#define blockDim (sdns_emu::tls().bdim)
#define gridDim (sdns_emu::tls().gdim)
inline void __syncthreads() { sdns_emu::tls().cta->all.arrive_and_wait(); }
inline void __syncwarp() { sdns_emu::Tls& t = sdns_emu::tls(); t.cta->warp[t.lin / 32]->arrive_and_wait(); }
template <typename Q> inline Q __shfl_sync(unsigned, Q v, int src) {
    static_assert(sizeof(Q) <= 8, "shuffle of at most 8 bytes");
    sdns_emu::Tls& t = sdns_emu::tls();
    unsigned long long* box = t.cta->mail.data() + (t.lin / 32) * 32;
    unsigned long long w = 0; std::memcpy(&w, &v, sizeof(Q));
    box[t.lin % 32] = w;
    t.cta->warp[t.lin / 32]->arrive_and_wait();
    const unsigned long long r = box[src & 31];
    t.cta->warp[t.lin / 32]->arrive_and_wait();
    Q out; std::memcpy(&out, &r, sizeof(Q)); return out;
}
template <typename Q> inline Q __shfl_xor_sync(unsigned m, Q v, int mask) { return __shfl_sync(m, v, (sdns_emu::tls().lin % 32) ^ mask); }

// ---- the slice of the CUDA runtime the library calls: memory is host memory, streams and events do nothing ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801 };
typedef struct sdns_emu_stream* cudaStream_t;
typedef struct sdns_emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaEventDisableTiming = 2,
       cudaStreamNonBlocking = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };
inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA runtime: not supported"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 2; return cudaSuccess; }      // "2 SMs"
template <class K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) std::memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
// graphs: a "capture" executes eagerly (everything is synchronous here) and a launch does nothing -- the library
// re-captures every step in this build
typedef struct sdns_emu_graph* cudaGraph_t;
typedef struct sdns_emu_graph_exec* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaSuccess; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaSuccess; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
// "IPC": the ranks of an emulated multi-GPU run are threads of one process, the handle carries the pointer
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof *h); std::memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

extern "C" inline __attribute__((visibility("default"), used)) void sdns_emu_set_skew(int microseconds) { sdns_emu::skew_us() = microseconds; }
