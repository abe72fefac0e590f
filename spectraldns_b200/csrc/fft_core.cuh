// fft_core.cuh -- register-resident mixed-radix Stockham line FFT for sm_100a.
//
// One FFT line of length N is owned by P = N/E threads; thread t holds the E elements
// x[q] = line[t + q*P].  Every stage (radix R | E, R in {8,4,2,3,5}) reads exactly that
// element set, so the global load pattern (t + q*P, coalesced over t or over the column
// index) is the same for every N, and the last stage leaves the result in natural order in
// the same register slots.  Between stages the line is exchanged through shared memory.
//
// This replaces the serial FFTW plans that shenfun / mpi4py-fft run underneath the
// reference's T.forward / T.backward (call sites solvers/NS.py:93,98,103,128,135).
#pragma once
#ifdef SDNS_HOST_SHIM          // tests/host: the same sources compiled by g++ against an emulation of the CUDA execution
#include "host_shim.h"         // model (kernel-logic tests without a GPU; never part of the product build)
#define SDNS_LAUNCH(kern, grid, block, smem, stream) sdns_emu::launcher(kern, grid, block, smem)
#define SDNS_DYN_SMEM(name) unsigned char* name = sdns_emu::dyn_smem()
#define SDNS_STATIC_SMEM(type, name, n) type* name = reinterpret_cast<type*>(sdns_emu::static_smem())
#else
#include <cuda_runtime.h>
#define SDNS_LAUNCH(kern, grid, block, smem, stream) kern<<<grid, block, smem, stream>>>
#define SDNS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SDNS_STATIC_SMEM(type, name, n) __shared__ type name[n]
#endif
#include <stdint.h>

namespace sdns {

template <typename T> struct C2;
template <> struct C2<float>  { typedef float2  type; };
template <> struct C2<double> { typedef double2 type; };

template <typename T> __device__ __forceinline__ typename C2<T>::type mk(T x, T y) {
    typename C2<T>::type r; r.x = x; r.y = y; return r;
}

// Two adjacent columns of an fp32 array in one register quad: a thread of a strided pass then moves 16 bytes per
// access, as an fp64 thread does, and every index, twiddle and shared-memory address serves two columns.
struct __align__(16) float2x2 { float2 a, b; };

// element traits: real type, scalar complex type (twiddles), columns per element
template <typename V> struct Elt;
template <> struct Elt<float2>   { typedef float T;  typedef float2 C;  static constexpr int cols = 1; };
template <> struct Elt<double2>  { typedef double T; typedef double2 C; static constexpr int cols = 1; };
template <> struct Elt<float2x2> { typedef float T;  typedef float2 C;  static constexpr int cols = 2; };

template <typename A, typename B> struct same_t { static constexpr bool v = false; };
template <typename A> struct same_t<A, A> { static constexpr bool v = true; };

// Complex arithmetic.  fp32 uses the packed f32x2 pipe of sm_100 (FADD2 / FMUL2 / FFMA2 on a (re, im) register
// pair; the lane swap and the sign of a multiplication by +-i fold into the operand modifiers), which halves the
// floating-point instruction count of the fp32 transforms.
template <typename V> __device__ __forceinline__ V cadd(V a, V b) {
    if constexpr (same_t<V, float2>::v) return __fadd2_rn(a, b);
    else if constexpr (same_t<V, float2x2>::v) { V r; r.a = __fadd2_rn(a.a, b.a); r.b = __fadd2_rn(a.b, b.b); return r; }
    else { a.x += b.x; a.y += b.y; return a; }
}
template <typename V> __device__ __forceinline__ V csub(V a, V b) {
    if constexpr (same_t<V, float2>::v) return __fadd2_rn(a, make_float2(-b.x, -b.y));
    else if constexpr (same_t<V, float2x2>::v) {
        V r; r.a = __fadd2_rn(a.a, make_float2(-b.a.x, -b.a.y)); r.b = __fadd2_rn(a.b, make_float2(-b.b.x, -b.b.y)); return r;
    } else { a.x -= b.x; a.y -= b.y; return a; }
}
// a * (-i) and a * (+i)
template <typename V> __device__ __forceinline__ V mul_mi(V a) {
    if constexpr (same_t<V, float2x2>::v) { V r; r.a = make_float2(a.a.y, -a.a.x); r.b = make_float2(a.b.y, -a.b.x); return r; }
    else { V r; r.x = a.y; r.y = -a.x; return r; }
}
template <typename V> __device__ __forceinline__ V mul_pi(V a) {
    if constexpr (same_t<V, float2x2>::v) { V r; r.a = make_float2(-a.a.y, a.a.x); r.b = make_float2(-a.b.y, a.b.x); return r; }
    else { V r; r.x = -a.y; r.y = a.x; return r; }
}
template <typename V> __device__ __forceinline__ V cneg(V a) {
    if constexpr (same_t<V, float2x2>::v) { V r; r.a = make_float2(-a.a.x, -a.a.y); r.b = make_float2(-a.b.x, -a.b.y); return r; }
    else { V r; r.x = -a.x; r.y = -a.y; return r; }
}
template <typename V> __device__ __forceinline__ V cconj(V a) { a.y = -a.y; return a; }
// rotate by -i for the forward transform (DIR=-1), +i for the backward one (DIR=+1)
template <int DIR, typename V> __device__ __forceinline__ V rot90(V a) {
    return DIR < 0 ? mul_mi(a) : mul_pi(a);
}
// a * s, s real
template <typename V> __device__ __forceinline__ V cscl(V a, typename Elt<V>::T s) {
    if constexpr (same_t<V, float2>::v) return __fmul2_rn(a, make_float2(s, s));
    else if constexpr (same_t<V, float2x2>::v) { V r; r.a = __fmul2_rn(a.a, make_float2(s, s)); r.b = __fmul2_rn(a.b, make_float2(s, s)); return r; }
    else { a.x *= s; a.y *= s; return a; }
}
// acc + x * c, c real
template <typename V> __device__ __forceinline__ V cfma(V acc, V x, typename Elt<V>::T c) {
    if constexpr (same_t<V, float2>::v) return __ffma2_rn(x, make_float2(c, c), acc);
    else if constexpr (same_t<V, float2x2>::v) {
        V r; r.a = __ffma2_rn(x.a, make_float2(c, c), acc.a); r.b = __ffma2_rn(x.b, make_float2(c, c), acc.b); return r;
    } else { acc.x += x.x * c; acc.y += x.y * c; return acc; }
}
// a * w, w one complex number (a twiddle)
__device__ __forceinline__ float2 cmul_f2(float2 a, float2 w) {
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
template <typename V> __device__ __forceinline__ V cmul(V a, typename Elt<V>::C w) {
    if constexpr (same_t<V, float2>::v) return cmul_f2(a, w);
    else if constexpr (same_t<V, float2x2>::v) { V r; r.a = cmul_f2(a.a, w); r.b = cmul_f2(a.b, w); return r; }
    else { V r; r.x = a.x * w.x - a.y * w.y; r.y = a.x * w.y + a.y * w.x; return r; }
}
// a * (wr + i wi) with compile-time constants
template <typename V> __device__ __forceinline__ V cmulc(V a, typename Elt<V>::T wr, typename Elt<V>::T wi) {
    if constexpr (same_t<V, double2>::v) { V r; r.x = a.x * wr - a.y * wi; r.y = a.x * wi + a.y * wr; return r; }
    else return cfma(cscl(a, wr), mul_pi(a), wi);
}

// ---------------------------------------------------------------------------------------
// small DFTs, natural-order in/out, exponent sign DIR (-1 forward, +1 backward)
// ---------------------------------------------------------------------------------------
template <int DIR, typename V>
__device__ __forceinline__ void dft2(V& a, V& b) {
    V t = a; a = cadd(t, b); b = csub(t, b);
}

template <int DIR, typename V>
__device__ __forceinline__ void dft4(V& a0, V& a1, V& a2, V& a3) {
    V t0 = cadd(a0, a2), t1 = csub(a0, a2);
    V t2 = cadd(a1, a3), t3 = rot90<DIR>(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

template <int DIR, typename T, typename V>
__device__ __forceinline__ void dft8(V& a0, V& a1, V& a2, V& a3, V& a4, V& a5, V& a6, V& a7) {
    const T h = (T)0.70710678118654752440084436210485L;
    // radix-2 split: evens (a0,a2,a4,a6), odds (a1,a3,a5,a7)
    V e0 = a0, e1 = a2, e2 = a4, e3 = a6;
    V o0 = a1, o1 = a3, o2 = a5, o3 = a7;
    dft4<DIR>(e0, e1, e2, e3);
    dft4<DIR>(o0, o1, o2, o3);
    // twiddles W8^k, k=1..3 with sign DIR:  W8^1 = (1 + DIR*i)/sqrt2 ; W8^2 = DIR*i ; W8^3 = (-1 + DIR*i)/sqrt2
    V w1 = cscl(cadd(o1, rot90<DIR>(o1)), h);
    V w3 = cscl(csub(rot90<DIR>(o3), o3), h);
    V w2 = rot90<DIR>(o2);
    a0 = cadd(e0, o0); a4 = csub(e0, o0);
    a1 = cadd(e1, w1); a5 = csub(e1, w1);
    a2 = cadd(e2, w2); a6 = csub(e2, w2);
    a3 = cadd(e3, w3); a7 = csub(e3, w3);
}

template <int DIR, typename T, typename V>
__device__ __forceinline__ void dft3(V& a0, V& a1, V& a2) {
    const T s = (T)0.86602540378443864676372317075294L;   // sin(pi/3)
    V t1 = cadd(a1, a2);
    V t2 = cfma(a0, t1, (T)-0.5);
    // forward: X1 = t2 - i*s*d ; backward: X1 = t2 + i*s*d
    V r = rot90<DIR>(cscl(csub(a1, a2), s));
    a0 = cadd(a0, t1);
    a1 = cadd(t2, r);
    a2 = csub(t2, r);
}

template <int DIR, typename T, typename V>
__device__ __forceinline__ void dft5(V& a0, V& a1, V& a2, V& a3, V& a4) {
    const T c1 = (T)0.30901699437494742410229341718282L;   // cos(2pi/5)
    const T c2 = (T)-0.80901699437494742410229341718282L;  // cos(4pi/5)
    const T s1 = (T)0.95105651629515357211643933337938L;   // sin(2pi/5)
    const T s2 = (T)0.58778525229247312916870595463907L;   // sin(4pi/5)
    V p1 = cadd(a1, a4), m1 = csub(a1, a4);
    V p2 = cadd(a2, a3), m2 = csub(a2, a3);
    V b1 = cfma(cfma(a0, p1, c1), p2, c2);
    V b2 = cfma(cfma(a0, p1, c2), p2, c1);
    V r1 = rot90<DIR>(cfma(cscl(m1, s1), m2, s2));
    V r2 = rot90<DIR>(cfma(cscl(m1, s2), m2, -s1));
    a0 = cadd(a0, cadd(p1, p2));
    a1 = cadd(b1, r1); a4 = csub(b1, r1);
    a2 = cadd(b2, r2); a3 = csub(b2, r2);
}

// multiply by W16^m (sign DIR), m a compile-time constant in 0..15
template <int DIR, int m, typename T, typename V>
__device__ __forceinline__ V mulw16(V a) {
    const T c = (T)0.92387953251128675612818318939679L;   // cos(pi/8)
    const T s = (T)0.38268343236508977172845998403040L;   // sin(pi/8)
    const T h = (T)0.70710678118654752440084436210485L;
    constexpr int mm = m & 15;
    if (mm == 0) return a;
    if (mm == 4) return rot90<DIR>(a);
    if (mm == 8) return cneg(a);
    if (mm == 12) return rot90<-DIR>(a);
    // W = cos(th) + i*DIR*sin(th), th = 2*pi*mm/16
    T wr, wi;
    if (mm == 1) { wr = c; wi = s; } else if (mm == 2) { wr = h; wi = h; } else if (mm == 3) { wr = s; wi = c; }
    else if (mm == 5) { wr = -s; wi = c; } else if (mm == 6) { wr = -h; wi = h; } else if (mm == 7) { wr = -c; wi = s; }
    else if (mm == 9) { wr = -c; wi = -s; } else if (mm == 10) { wr = -h; wi = -h; } else if (mm == 11) { wr = -s; wi = -c; }
    else if (mm == 13) { wr = s; wi = -c; } else if (mm == 14) { wr = h; wi = -h; } else { wr = c; wi = -s; }
    if (DIR < 0) wi = -wi;
    return cmulc(a, wr, wi);
}

// 16-point DFT as 4 x 4: radix-4 over n2 (stride 4), twiddle W16^(n1*k2), radix-4 over n1, transpose
template <int DIR, typename T, typename V>
__device__ __forceinline__ void dft16(V& a0, V& a1, V& a2, V& a3, V& a4, V& a5, V& a6, V& a7,
                                      V& a8, V& a9, V& a10, V& a11, V& a12, V& a13, V& a14, V& a15) {
    dft4<DIR>(a0, a4, a8, a12);      // n1 = 0: slots (0,4,8,12) <- k2 = 0..3
    dft4<DIR>(a1, a5, a9, a13);      // n1 = 1
    dft4<DIR>(a2, a6, a10, a14);     // n1 = 2
    dft4<DIR>(a3, a7, a11, a15);     // n1 = 3
    // slot n1 + 4*k2 holds B[n1][k2]; multiply by W16^(n1*k2)
    a5 = mulw16<DIR, 1, T>(a5);  a9 = mulw16<DIR, 2, T>(a9);   a13 = mulw16<DIR, 3, T>(a13);
    a6 = mulw16<DIR, 2, T>(a6);  a10 = mulw16<DIR, 4, T>(a10); a14 = mulw16<DIR, 6, T>(a14);
    a7 = mulw16<DIR, 3, T>(a7);  a11 = mulw16<DIR, 6, T>(a11); a15 = mulw16<DIR, 9, T>(a15);
    dft4<DIR>(a0, a1, a2, a3);       // k2 = 0: slots (0,1,2,3) <- k1 = 0..3  => X[4*k1 + 0]
    dft4<DIR>(a4, a5, a6, a7);       // k2 = 1                                 => X[4*k1 + 1]
    dft4<DIR>(a8, a9, a10, a11);     // k2 = 2
    dft4<DIR>(a12, a13, a14, a15);   // k2 = 3
    // slot k1 + 4*k2 holds X[4*k1 + k2]: transpose to natural order
    V t;
    t = a1; a1 = a4; a4 = t;      t = a2; a2 = a8; a8 = t;      t = a3; a3 = a12; a12 = t;
    t = a6; a6 = a9; a9 = t;      t = a7; a7 = a13; a13 = t;    t = a11; a11 = a14; a14 = t;
}

// cos / sin of 2 pi m / 24 (the twiddles inside the composite radix-12 and radix-24 butterflies)
template <typename T> __host__ __device__ constexpr T cos24(int m) {
    switch (((m % 24) + 24) % 24) {
        case 0: return (T)1; case 1: case 23: return (T)0.96592582628906828674974319972890L;
        case 2: case 22: return (T)0.86602540378443864676372317075294L; case 3: case 21: return (T)0.70710678118654752440084436210485L;
        case 4: case 20: return (T)0.5; case 5: case 19: return (T)0.25881904510252076234889883762405L;
        case 6: case 18: return (T)0; case 7: case 17: return (T)-0.25881904510252076234889883762405L;
        case 8: case 16: return (T)-0.5; case 9: case 15: return (T)-0.70710678118654752440084436210485L;
        case 10: case 14: return (T)-0.86602540378443864676372317075294L; case 11: case 13: return (T)-0.96592582628906828674974319972890L;
        default: return (T)-1;      // 12
    }
}
template <typename T> __host__ __device__ constexpr T sin24(int m) { return cos24<T>(m - 6); }

// R-point DFT with R = R1 * 3 (R1 = 4 or 8) entirely in registers, natural order in and out: Cooley-Tukey with
// n = 3 n1 + n2, k = k1 + R1 k2 -- three DFT-R1 over n1, the twiddles W_R^(n2 k1), R1 DFT-3 over n2.  Lets the lengths
// 3 * 2^k (the 3/2-rule paddings) take a factor 12 or 24 per stage instead of 4 / 8 and a separate factor 3, i.e. one
// shared-memory exchange fewer per line.
template <int DIR, int R1, typename T, typename V>
__device__ __forceinline__ void dft_x3(V (&a)[R1 * 3]) {
    constexpr int R = R1 * 3;
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2) {
        if constexpr (R1 == 8) dft8<DIR, T>(a[n2], a[3 + n2], a[6 + n2], a[9 + n2], a[12 + n2], a[15 + n2], a[18 + n2], a[21 + n2]);
        else dft4<DIR>(a[n2], a[3 + n2], a[6 + n2], a[9 + n2]);
    }
    // slot 3 k1 + n2 now holds Y[k1][n2]
#pragma unroll
    for (int k1 = 1; k1 < R1; ++k1)
#pragma unroll
        for (int n2 = 1; n2 < 3; ++n2) {
            const int m = (24 / R) * n2 * k1;                       // W_R^(n2 k1) = W_24^m
            a[3 * k1 + n2] = cmulc(a[3 * k1 + n2], cos24<T>(m), (DIR < 0 ? -1 : 1) * sin24<T>(m));
        }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) dft3<DIR, T>(a[3 * k1], a[3 * k1 + 1], a[3 * k1 + 2]);
    // slot 3 k1 + k2 holds X[k1 + R1 k2]
    V b[R];
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) b[k1 + R1 * k2] = a[3 * k1 + k2];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = b[i];
}

// ---------------------------------------------------------------------------------------
// radix plan: largest radix in {24,16,12,8,4,2,3,5} dividing both what is left of N and E
// ---------------------------------------------------------------------------------------
__host__ __device__ constexpr int pick_radix(int rem, int E) {
    return
#ifndef SDNS_NO_RADIX_X3
           (rem % 24 == 0 && E % 24 == 0) ? 24 :
           (rem % 12 == 0 && E % 12 == 0 && E % 8 != 0) ? 12 :
#endif
           (rem % 16 == 0 && E % 16 == 0) ? 16 :
           (rem % 8 == 0 && E % 8 == 0) ? 8 :
           (rem % 4 == 0 && E % 4 == 0) ? 4 :
           (rem % 2 == 0 && E % 2 == 0) ? 2 :
           (rem % 3 == 0 && E % 3 == 0) ? 3 :
           (rem % 5 == 0 && E % 5 == 0) ? 5 : 0;
}
__host__ __device__ constexpr bool plan_ok(int N, int E) {
    int ns = 1;
    while (ns < N) { int r = pick_radix(N / ns, E); if (r == 0) return false; ns *= r; }
    return ns == N && N % E == 0;
}
__host__ __device__ constexpr int num_stages(int N, int E) {
    int ns = 1, s = 0;
    while (ns < N) { int r = pick_radix(N / ns, E); if (r == 0) return -1; ns *= r; ++s; }
    return s;
}

// Shared-memory addressing of one line during an exchange.
//   contiguous-line kernels: addr = base + idx + idx/PADW   (PADW elements = 128 bytes)
//   strided (column-tile) kernels: addr = idx*TCOLS + column  (a quarter/half warp touches
//   TCOLS contiguous elements = 128 bytes: conflict free for every stage)
template <int MUL, int PADW>
struct SmemLine {
    int base;
    __device__ __forceinline__ int operator()(int idx) const {
        if constexpr (PADW > 0) return base + idx + idx / PADW;
        else return base + idx * MUL;
    }
};

// SYNC: 0 = __syncthreads (line spans warps), 1 = __syncwarp (line lives inside one warp)
template <int SYNC> __device__ __forceinline__ void line_sync() {
    if (SYNC == 1) __syncwarp(); else __syncthreads();
}

template <typename T, int N, int E, int DIR, int Ns, int R, typename V>
__device__ __forceinline__ void fft_stage(V (&x)[E], int t, const typename Elt<V>::C* __restrict__ tw) {
    typedef typename Elt<V>::C W;
    constexpr int P = N / E;
    constexpr int NB = E / R;          // butterflies per thread
    constexpr int TS = N / (Ns * R);   // twiddle table stride
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        if (Ns > 1) {
            const int jm = (t + m * P) % Ns;
#ifdef SDNS_TW_TABLE_ALL
#pragma unroll
            for (int k = 1; k < R; ++k) {
                W w = __ldg(&tw[(jm * k) * TS]);
                if (DIR > 0) w.y = -w.y;
                x[m + k * NB] = cmul(x[m + k * NB], w);
            }
#else
            // one table load per butterfly; the powers w^2..w^(R-1) by multiplication (the L1/LSU path,
            // not the FP pipe, is the busy unit of these kernels)
            W w1 = __ldg(&tw[jm * TS]);
            if (DIR > 0) w1.y = -w1.y;
            x[m + NB] = cmul(x[m + NB], w1);
            // (the composite radices 12 / 24 only ever form the FIRST stage of the compiled lengths, which has no twiddles)
            static_assert(!(R == 12 || R == 24) || Ns == 1, "composite radix beyond the first stage");
            if (R > 2) {
                const W w2 = cmul(w1, w1);
                x[m + 2 * NB] = cmul(x[m + 2 * NB], w2);
                if (R > 3) {
                    const W w3 = cmul(w2, w1);
                    x[m + 3 * NB] = cmul(x[m + 3 * NB], w3);
                    if (R > 4) {
                        const W w4 = cmul(w2, w2);
                        x[m + 4 * NB] = cmul(x[m + 4 * NB], w4);
                        if (R > 5) {
                            const W w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w4, w3);
                            x[m + 5 * NB] = cmul(x[m + 5 * NB], w5);
                            x[m + 6 * NB] = cmul(x[m + 6 * NB], w6);
                            x[m + 7 * NB] = cmul(x[m + 7 * NB], w7);
                            if (R > 8) {
                                const W w8 = cmul(w4, w4);
                                x[m + 8 * NB] = cmul(x[m + 8 * NB], w8);
                                x[m + 9 * NB] = cmul(x[m + 9 * NB], cmul(w8, w1));
                                x[m + 10 * NB] = cmul(x[m + 10 * NB], cmul(w5, w5));
                                x[m + 11 * NB] = cmul(x[m + 11 * NB], cmul(w8, w3));
                                x[m + 12 * NB] = cmul(x[m + 12 * NB], cmul(w6, w6));
                                x[m + 13 * NB] = cmul(x[m + 13 * NB], cmul(w8, w5));
                                x[m + 14 * NB] = cmul(x[m + 14 * NB], cmul(w7, w7));
                                x[m + 15 * NB] = cmul(x[m + 15 * NB], cmul(w8, w7));
                            }
                        }
                    }
                }
            }
#endif
        }
        if constexpr (R == 12 || R == 24) {
            V a[R];
#pragma unroll
            for (int k = 0; k < R; ++k) a[k] = x[m + k * NB];
            dft_x3<DIR, R / 3, T>(a);
#pragma unroll
            for (int k = 0; k < R; ++k) x[m + k * NB] = a[k];
        } else
        if (R == 2) dft2<DIR>(x[m], x[m + NB]);
        else if (R == 4) dft4<DIR>(x[m], x[m + NB], x[m + 2 * NB], x[m + 3 * NB]);
        else if (R == 8) dft8<DIR, T>(x[m], x[m + NB], x[m + 2 * NB], x[m + 3 * NB],
                                      x[m + 4 * NB], x[m + 5 * NB], x[m + 6 * NB], x[m + 7 * NB]);
        else if (R == 16) dft16<DIR, T>(x[m], x[m + NB], x[m + 2 * NB], x[m + 3 * NB], x[m + 4 * NB], x[m + 5 * NB],
                                        x[m + 6 * NB], x[m + 7 * NB], x[m + 8 * NB], x[m + 9 * NB], x[m + 10 * NB],
                                        x[m + 11 * NB], x[m + 12 * NB], x[m + 13 * NB], x[m + 14 * NB], x[m + 15 * NB]);
        else if (R == 3) dft3<DIR, T>(x[m], x[m + NB], x[m + 2 * NB]);
        else if (R == 5) dft5<DIR, T>(x[m], x[m + NB], x[m + 2 * NB], x[m + 3 * NB], x[m + 4 * NB]);
    }
}

template <typename T, int N, int E, int Ns, int R, typename V, typename SM>
__device__ __forceinline__ void fft_scatter(const V (&x)[E], int t, V* sm, const SM& map) {
    constexpr int P = N / E;
    constexpr int NB = E / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * P;
        const int o = (j / Ns) * (Ns * R) + (j % Ns);
#pragma unroll
        for (int k = 0; k < R; ++k) sm[map(o + k * Ns)] = x[m + k * NB];
    }
}

template <typename T, int N, int E, typename V, typename SM>
__device__ __forceinline__ void fft_gather(V (&x)[E], int t, const V* sm, const SM& map) {
    constexpr int P = N / E;
#pragma unroll
    for (int q = 0; q < E; ++q) x[q] = sm[map(t + q * P)];
}

// Full line transform.  sm points at this CTA's exchange region; `map` places this line in it.
// NBUF = 2: ping-pong between two regions `bufstride` elements apart (one sync per exchange);
// NBUF = 1: single region, two syncs per exchange.  `phase` must be CTA-uniform.
template <typename T, int N, int E, int DIR, int SYNC, int NBUF, int Ns, typename V, typename SM>
__device__ __forceinline__ void fft_stages(V (&x)[E], int t, const typename Elt<V>::C* __restrict__ tw,
                                           V* sm, const SM& map, int bufstride, int& phase) {
    if constexpr (Ns < N) {
        constexpr int R = pick_radix(N / Ns, E);
        static_assert(R > 0, "unsupported FFT length for this E");
        fft_stage<T, N, E, DIR, Ns, R>(x, t, tw);
        if constexpr (Ns * R < N) {
            V* b = sm + (NBUF == 2 ? phase * bufstride : 0);
            if (NBUF == 1) line_sync<SYNC>();
            fft_scatter<T, N, E, Ns, R>(x, t, b, map);
            line_sync<SYNC>();
            fft_gather<T, N, E>(x, t, b, map);
            if (NBUF == 2) phase ^= 1;
            fft_stages<T, N, E, DIR, SYNC, NBUF, Ns * R>(x, t, tw, sm, map, bufstride, phase);
        }
    }
}

template <typename T, int N, int E, int DIR, int SYNC, int NBUF, typename V, typename SM>
__device__ __forceinline__ void fft_line(V (&x)[E], int t, const typename Elt<V>::C* __restrict__ tw,
                                         V* sm, const SM& map, int bufstride, int& phase) {
    fft_stages<T, N, E, DIR, SYNC, NBUF, 1>(x, t, tw, sm, map, bufstride, phase);
}

// wavenumber (integer) of memory index i on a c2c axis of length n (numpy.fft.fftfreq order)
__device__ __forceinline__ int wavenum(int i, int n) { return i < (n + 1) / 2 ? i : i - n; }

}  // namespace sdns
