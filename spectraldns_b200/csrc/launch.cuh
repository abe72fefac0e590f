// launch.cuh -- size dispatch for the pass kernels.  Every (family, precision) pair is its own
// translation unit (inst.cu compiled with -DSDNS_FAMILY=.. -DSDNS_REAL=..), so the sm_100a
// build parallelises over the host cores.
#pragma once
#include "passes.cuh"

namespace sdns {

// transform lengths with a compiled kernel: 2^k (8..2048), 3*2^k (12..3072, the 3/2-rule lengths), and 60 / 90
// (demo/Isotropic.py's own default grid and its 3/2 padding: radices 2*2*3*5 and 2*3*3*5, 30 elements per thread)
// 60 / 90 are compiled for the NS / VV Vortex path and the plain transforms only (SDNS_SIZES_5): their kernels are
// large (30 elements per thread) and MHD / the other NS convection forms report "no kernel for length" there.
#ifndef SDNS_SIZES
#define SDNS_SIZES(X) X(8) X(12) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) \
                      X(24) X(48) X(96) X(192) X(384) X(768) X(1536) X(3072)
#define SDNS_SIZES_5(X) X(60) X(90)
#endif
#ifndef SDNS_SIZES_5
#define SDNS_SIZES_5(X)
#endif

enum Family {
    FAM_PLAIN_FWD = 0, FAM_PLAIN_BWD = 1, FAM_NS_B0 = 2, FAM_VV_B0 = 3, FAM_NS_F0 = 4, FAM_VV_F0 = 5,
    FAM_MHD_F0 = 6, FAM_Z_C2R = 7, FAM_Z_R2C = 8, FAM_Z_CROSS = 9, FAM_Z_MHD = 10,
    FAM_NS_GRAD_B0 = 11, FAM_Z_DOT = 12, FAM_Z_UU = 13, FAM_NSDIV_F0 = 14, FAM_Z_NS2D = 15, FAM_Z_BQ2D = 16, FAM_COUNT = 17
};

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// launch geometry of the strided kernels
// elements per thread: keep a line inside one warp (P = N/E <= 32) while the register budget
// allows it (fp32: E <= 32, fp64: E <= 16 for single-field kernels), else 64 or 128 threads per line
constexpr int pick_E(int N, int emax) {
    if (N % 5 == 0) return 30;              // a factor 5: every stage radix (2, 3, 5) must divide E
    const int b = (N % 3 == 0) ? 12 : 8;
    int e = b;
    while (N / e > 32 && 2 * e <= emax && plan_ok(N, 2 * e)) e *= 2;
    while (N / e > 128 && plan_ok(N, 2 * e)) e *= 2;
    return e;
}

// fp64 strided kernels: P <= 64 threads per line (E up to 16) and CTAs of at most 256 threads, so that
// three or four CTAs are resident per SM and their load / transform / store phases overlap
constexpr int pick_E64(int N) {
    if (N % 5 == 0) return 30;
    const int b = (N % 3 == 0) ? 12 : 8;
#ifndef SDNS_NO_RADIX16
    if (b == 8 && N >= 256 && N % 256 == 0) {           // 16 x 16 (x 2,4,8): one exchange fewer than 8 x 8 x 4
        int e = 16;
        while (N / e > 64 && plan_ok(N, 2 * e) && 2 * e <= 32) e *= 2;
        return e;
    }
#endif
    int e = b;
    while (N / e > 64 && 2 * e <= 16 + (b == 12 ? 8 : 0) && plan_ok(N, 2 * e)) e *= 2;
    while (N / e > 128 && plan_ok(N, 2 * e)) e *= 2;
    return e;
}

template <typename T, int N, int MODE>
struct SCfg {
    static constexpr bool heavy = (MODE == S_NS_F0 || MODE == S_VV_F0);   // park two fields in smem
    static constexpr bool b0e = (MODE == S_NS_B0 || MODE == S_VV_B0 || MODE == S_NS_GRAD_B0);
#ifndef SDNS_F32_EMAX
#define SDNS_F32_EMAX 24
#endif
#ifndef SDNS_F32_EMAX_HEAVY
#define SDNS_F32_EMAX_HEAVY 16
#endif
#ifndef SDNS_F32_MAXT
#define SDNS_F32_MAXT 256
#endif
#ifndef SDNS_F32_MAXT_HEAVY
#define SDNS_F32_MAXT_HEAVY 512
#endif
    static constexpr int E = sizeof(T) == 8 ? ((heavy || b0e) ? pick_E(N, 8) : pick_E64(N))
                                            : pick_E(N, heavy ? SDNS_F32_EMAX_HEAVY : SDNS_F32_EMAX);
    static constexpr int P = N / E;
    // the passes that store into peer GPUs (B0 family) keep 128-byte rows: NVLink likes the larger packets
    static constexpr bool b0m = (MODE == S_NS_B0 || MODE == S_VV_B0 || MODE == S_NS_GRAD_B0);
#ifndef SDNS_B0_MAXT
#define SDNS_B0_MAXT 512
#endif
#ifndef SDNS_F64_MAXT_HEAVY
#define SDNS_F64_MAXT_HEAVY 256
#endif
    static constexpr int maxThreads = sizeof(T) == 8 ? (b0m ? SDNS_B0_MAXT : (heavy ? SDNS_F64_MAXT_HEAVY : 256)) : (heavy ? SDNS_F32_MAXT_HEAVY : SDNS_F32_MAXT);
    static constexpr int TCfull = 128 / (2 * (int)sizeof(T));
    // widest tile whose buffers (the exchange buffer, plus two parked fields in the epilogue kernels) fit in shared memory
    static constexpr int tc_fit(int tc) {
        while (tc > 1 && (size_t)N * tc * 2 * sizeof(T) * (heavy ? 3 : 1) > 200 * 1024) tc /= 2;
        return tc;
    }
    static constexpr int TC = tc_fit(cmin(TCfull, cmax(1, maxThreads / P)));
    static constexpr size_t bytes1 = (size_t)N * TC * 2 * sizeof(T);
#ifdef SDNS_STRIDED_NBUF
    static constexpr int NBUF = SDNS_STRIDED_NBUF;
#else
    static constexpr int NBUF = (heavy || num_stages(N, E) <= 2) ? 1 : ((2 * bytes1 <= 100 * 1024) ? 2 : 1);
#endif
    static constexpr size_t smem = bytes1 * (NBUF + (heavy ? 2 : 0));
    // resident CTAs per SM the register allocator must allow: what shared memory and the thread
    // count permit, at most 4 (2 for the epilogue kernels)
    static constexpr int bySmem = (int)((224 * 1024) / (smem + 1024));
    static constexpr int byThreads = 2048 / (P * TC);
    static constexpr bool b0 = (MODE == S_NS_B0 || MODE == S_VV_B0 || MODE == S_NS_GRAD_B0);
#ifndef SDNS_PLAIN64_REGS
#define SDNS_PLAIN64_REGS 128
#endif
    static constexpr int needRegs = sizeof(T) == 8 ? (heavy ? 128 : (E > 12 ? (b0 ? 168 : SDNS_PLAIN64_REGS) : (b0 ? 96 : 80)))
                                                   : (heavy ? (E > 12 ? 128 : 100) : (E > 16 ? 80 : (b0 ? 72 : 64)));
    static constexpr int byRegs = 65536 / (P * TC * needRegs);
    static constexpr int minBlocks = cmax(1, cmin(cmin(cmin(bySmem, byThreads), byRegs), heavy ? 2 : 4));
};

// field-parallel F0 (f0x_kernel): 3 thread groups; shared memory = 3 exchange buffers that double as parking
constexpr int fx_blocks(int P, int tc, int csize, int N, int regs) {
    const int threads = 3 * P * tc;
    if (threads > 1024) return 0;
    const long long smem = 3LL * N * tc * csize;
    if (smem > 200 * 1024) return 0;
    const int bs = (int)((216 * 1024) / (smem + 1024)), br = 65536 / (threads * regs), bt = 2048 / threads;
    return cmin(cmin(bs, br), cmin(bt, 4));
}
// tile width: the widest one (rows of >= 64 bytes) that still lets two CTAs share an SM, else the widest that fits
constexpr int fx_tc(int P, int tcfull, int csize, int N, int regs) {
    const int tmin = 64 / csize;
    for (int tc = tcfull; tc >= tmin; tc /= 2) if (fx_blocks(P, tc, csize, N, regs) >= 2) return tc;
    for (int tc = tcfull; tc >= 1; tc /= 2) if (fx_blocks(P, tc, csize, N, regs) >= 1) return tc;
    return 0;
}

template <typename T, int N>
struct FXCfg {
#ifdef SDNS_F0_E16
    static constexpr int E = sizeof(T) == 8 ? pick_E64(N) : pick_E(N, 16);
#else
    static constexpr int E = sizeof(T) == 8 ? pick_E(N, 8) : pick_E(N, 16);
#endif
    static constexpr int P = N / E;
    static constexpr int csize = 2 * (int)sizeof(T);
    static constexpr int needRegs = sizeof(T) == 8 ? (E > 12 ? 128 : 84) : (E > 12 ? 84 : 64);
#ifdef SDNS_FX_TC
    static constexpr int TC = fx_blocks(P, SDNS_FX_TC, csize, N, needRegs) >= 1 ? SDNS_FX_TC : fx_tc(P, 128 / csize, csize, N, needRegs);
#else
    static constexpr int TC = fx_tc(P, 128 / csize, csize, N, needRegs);
#endif
    static constexpr bool ok = TC > 0 && plan_ok(N, E) && sizeof(T) == 8;   // fp32: the register version is faster
    static constexpr size_t smem = (size_t)3 * N * (TC > 0 ? TC : 1) * csize;
    static constexpr int threads = 3 * P * (TC > 0 ? TC : 1);
    static constexpr int minBlocks = cmax(1, fx_blocks(P, TC > 0 ? TC : 1, csize, N, needRegs));
};

// field-parallel B0 (b0x_kernel): 3 thread groups; shared memory = 3 exchange buffers + 3 slot arrays.  Used when a
// tile of >= 64-byte rows fits (fp64: N <= 512, fp32: N <= 1024); the other lengths keep strided_kernel's B0 branch.
constexpr int bx_tc(int P, int tcfull, int csize, int N) {
    for (int tc = tcfull; tc * csize >= 64; tc /= 2)
        if (3 * P * tc <= 768 && 6LL * N * tc * csize <= 200 * 1024) return tc;
    return 0;
}
template <typename T, int N>
struct BXCfg {
    static constexpr int E = sizeof(T) == 8 ? pick_E(N, 8) : pick_E(N, 16);
    static constexpr int P = N / E;
    static constexpr int csize = 2 * (int)sizeof(T);
    static constexpr int TC = bx_tc(P, 128 / csize, csize, N);
#ifndef SDNS_B0X          // opt-in experiment: measured on B200, not faster than strided_kernel's B0 branch
    static constexpr bool ok = false;
#else
    static constexpr bool ok = TC > 0 && plan_ok(N, E) && N % 5 != 0;
#endif
    static constexpr size_t smem = (size_t)6 * N * (TC > 0 ? TC : 1) * csize;
    static constexpr int threads = 3 * P * (TC > 0 ? TC : 1);
    static constexpr int bySmem = (int)((224 * 1024) / (smem + 1024));
    static constexpr int needRegs = sizeof(T) == 8 ? 84 : 64;
    static constexpr int minBlocks = cmax(1, cmin(cmin(bySmem, 2048 / threads), cmin(65536 / (threads * needRegs), 4)));
};

// MHD epilogue: six accumulators per thread -> fewer elements per thread
template <typename T, int N>
struct MCfg {
#ifndef SDNS_MHD_F32_E3
#define SDNS_MHD_F32_E3 12      // fp32, 3*2^k lengths: elements per thread (six accumulators of E complex values per thread)
#endif
    static constexpr int E = (N % 5 == 0) ? 30 : (N % 3 == 0) ? (sizeof(T) == 8 ? 6 : SDNS_MHD_F32_E3) : (sizeof(T) == 8 ? 4 : 8);
    static constexpr int P = N / E;
    static constexpr int maxThreads = 512;
    static constexpr int TCfull = 128 / (2 * (int)sizeof(T));
    static constexpr int TC = cmin(TCfull, cmax(1, maxThreads / P));
    static constexpr size_t bytes1 = (size_t)N * TC * 2 * sizeof(T);
    static constexpr int NBUF = (2 * bytes1 <= 100 * 1024) ? 2 : 1;
    static constexpr size_t smem = bytes1 * NBUF;
};

template <typename T, int M, int MODE>
struct ZCfg {
    static constexpr bool park = (MODE == Z_CROSS);     // two real-space pairs parked in smem
    static constexpr int E = pick_E(M, sizeof(T) == 8 ? ((MODE == Z_C2R || MODE == Z_R2C) ? 16 : 12) : (MODE == Z_MHD ? 16 : 32));
    static constexpr int P = M / E;
    static constexpr int LPC = cmax(1, 128 / P);
    static constexpr int SYNC = (P <= 32 && 32 % P == 0) ? 1 : 0;     // __syncwarp only if no line straddles two warps
    static constexpr int NBUF = park ? 1 : 2;
    static constexpr int PADW = 128 / (2 * (int)sizeof(T));
    static constexpr int LP = M + M / PADW + 1;
    static constexpr size_t smem = ((size_t)LP * LPC * NBUF + (park ? (size_t)2 * E * P * LPC : 0)) * 2 * sizeof(T);
    static constexpr int needRegs = (sizeof(T) == 8 ? (E > 12 ? 200 : 128) : (E > 16 ? 128 : 96)) * ((MODE == Z_MHD || MODE == Z_DOT || MODE == Z_NS2D || MODE == Z_BQ2D) ? 2 : 1);
    static constexpr int minBlocks = cmax(1, cmin(cmin(cmin((int)((224 * 1024) / (smem + 1024)), 2048 / (P * LPC)),
                                                       65536 / (P * LPC * needRegs)), 4));
};

// warp-per-line fused z kernel (zx_kernel): available when M/E == 32
template <typename T, int M>
struct ZXCfg {
    static constexpr int E = M / 32;
#ifndef SDNS_ZX_F32_EMAX
#define SDNS_ZX_F32_EMAX 32
#endif
    static constexpr bool ok = (M % 32 == 0) && (E == 8 || E == 12 || E == 16 || E == 24 || E == 32) && plan_ok(M, E)
                               && (sizeof(T) == 4 ? E <= SDNS_ZX_F32_EMAX : E <= 12);
    static constexpr int LPC = 4;
    static constexpr int PADW = 128 / (2 * (int)sizeof(T));
    static constexpr int LP = M + M / PADW + 1;
    static constexpr size_t smem = ((size_t)LP * LPC + (size_t)2 * E * 32 * LPC) * 2 * sizeof(T);
    static constexpr int QN3 = (M / 3 + 2 + 31) / 32;        // covers the 2/3-rule and 3/2-rule mode counts
    static constexpr int QN2 = E / 2 + 1;                    // all M/2+1 modes
#ifndef SDNS_ZX_E8_REGS64
#define SDNS_ZX_E8_REGS64 168
#endif
    static constexpr int needRegs = sizeof(T) == 8 ? (E > 8 ? 255 : SDNS_ZX_E8_REGS64) : (E > 16 ? 255 : (E > 8 ? 168 : 128));
    static constexpr int minBlocks = cmax(1, cmin(cmin((int)((216 * 1024) / (smem + 1024)), 16), 65536 / (128 * needRegs)));
};

// zx_kernel with bulk-async input staging (zb_kernel): same lengths; LPC warps per CTA, NST stages per warp
#ifndef SDNS_ZB_NST
#define SDNS_ZB_NST 2
#endif
#ifndef SDNS_ZB_LPC
#define SDNS_ZB_LPC 2
#endif
#ifndef SDNS_ZB_REGS64
#define SDNS_ZB_REGS64 128
#endif
#ifndef SDNS_ZB_REGS32
#define SDNS_ZB_REGS32 128
#endif
template <typename T, int M>
struct ZBCfg {
    static constexpr int E = ZXCfg<T, M>::E;
    static constexpr bool ok = ZXCfg<T, M>::ok;
    static constexpr int LPC = SDNS_ZB_LPC, NST = SDNS_ZB_NST;
    static constexpr int PADW = 128 / (2 * (int)sizeof(T));
    static constexpr int LP = (M + M / PADW + 2) & ~1;
    static constexpr size_t smem_q(int qn) { return 128 + (size_t)LPC * (LP + 2 * E * 32 + NST * 6 * 32 * qn) * 2 * sizeof(T); }
    static constexpr int needRegs = sizeof(T) == 8 ? (E > 8 ? 200 : SDNS_ZB_REGS64) : (E > 16 ? 200 : (E > 8 ? 168 : SDNS_ZB_REGS32));
    static constexpr int minBlocks_q(int qn) {
        return cmax(1, cmin(cmin((int)((224 * 1024) / (smem_q(qn) + 1024)), 32), 65536 / (32 * LPC * needRegs)));
    }
};

// CTA-per-line fused z kernel (zy_kernel): two or four warps per line, for the lengths zx_kernel cannot hold in
// one warp's registers
#ifndef SDNS_ZY_E4_REGS
#define SDNS_ZY_E4_REGS 96
#endif
constexpr int zy_E(int M) {
#ifdef SDNS_ZY_E4
    if (M % 4 == 0 && (M / 4 == 64 || M / 4 == 128) && plan_ok(M, 4)) return 4;
#endif
    return (M % 8 == 0 && M / 8 <= 128 && plan_ok(M, 8)) ? 8 :
           (M % 12 == 0 && M / 12 <= 128 && plan_ok(M, 12)) ? 12 :
           (M % 16 == 0 && M / 16 <= 128 && plan_ok(M, 16)) ? 16 :
           (M % 24 == 0 && M / 24 <= 128 && plan_ok(M, 24)) ? 24 : 0;
}
template <typename T, int M>
struct ZYCfg {
    static constexpr int E = zy_E(M) > 0 ? zy_E(M) : 8;
    static constexpr int P = M / E;
#ifdef SDNS_ZY_E4
    static constexpr bool ok = zy_E(M) > 0 && (P == 64 || P == 128) && (!ZXCfg<T, M>::ok || (sizeof(T) == 8 && E == 4));
#else
    static constexpr bool ok = zy_E(M) > 0 && (P == 64 || P == 128) && !ZXCfg<T, M>::ok;
#endif
    static constexpr int PADW = 128 / (2 * (int)sizeof(T));
    static constexpr int LP = M + M / PADW + 1;
    static constexpr int QN3 = (M / 3 + 2 + P - 1) / P;       // covers the 2/3-rule and 3/2-rule mode counts
    static constexpr int QN2 = (M / 2 + 1 + P - 1) / P;       // all M/2+1 modes
    static constexpr size_t smem_q(int qn) { return ((size_t)(LP > 2 * qn * P ? LP : 2 * qn * P) + (size_t)2 * E * P) * 2 * sizeof(T); }
#ifndef SDNS_ZY_E8_REGS64
#define SDNS_ZY_E8_REGS64 168
#endif
    static constexpr int needRegs = sizeof(T) == 8 ? (E > 12 ? 255 : (E > 8 ? 224 : (E > 4 ? SDNS_ZY_E8_REGS64 : SDNS_ZY_E4_REGS))) : (E > 16 ? 224 : (E > 12 ? 168 : 128));
    static constexpr int minBlocks_q(int qn) {
        return cmax(1, cmin(cmin((int)((216 * 1024) / (smem_q(qn) + 1024)), 16), 65536 / (P * needRegs)));
    }
};

template <typename K>
inline cudaError_t set_smem(K kern, size_t smem) {
    if (smem > 48 * 1024)
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaSuccess;
}

// returns cudaError_t as int, or -1000 when the length has no compiled kernel
template <typename T> int launch_family(int family, int n, const void* args, cudaStream_t st);

// entry points defined by the instantiation units (one per family and precision)
#define SDNS_DECL(fam) \
    int sdns_launch_##fam##_f32(int n, const void* args, cudaStream_t st); \
    int sdns_launch_##fam##_f64(int n, const void* args, cudaStream_t st);
SDNS_DECL(0) SDNS_DECL(1) SDNS_DECL(2) SDNS_DECL(3) SDNS_DECL(4) SDNS_DECL(5)
SDNS_DECL(6) SDNS_DECL(7) SDNS_DECL(8) SDNS_DECL(9) SDNS_DECL(10)
SDNS_DECL(11) SDNS_DECL(12) SDNS_DECL(13) SDNS_DECL(14) SDNS_DECL(15) SDNS_DECL(16)
#undef SDNS_DECL

}  // namespace sdns
