// host_shim.h -- TEST INFRASTRUCTURE ONLY: lets g++ compile fft_core.cuh for the CPU, so that the radix plans,
// butterflies, twiddle indexing and exchange maps of the product's FFT core are checked without a GPU
// (tests/test_fft_core_host.py).  Never included in the CUDA build.
#pragma once
#include <cmath>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __align__(n) alignas(n)
struct float2 { float x, y; };
struct double2 { double x, y; };
inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
// the packed f32x2 intrinsics of sm_100: lane-wise, round to nearest
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)); }
template <typename Q> inline Q __ldg(const Q* p) { return *p; }
inline void __syncwarp() {}
inline void __syncthreads() {}
