// launch.cuh -- size dispatch for the pass kernels.  Every (family, precision) pair is its own
// translation unit (inst.cu compiled with -DSDNS_FAMILY=.. -DSDNS_REAL=..), so the sm_100a
// build parallelises over the host cores.
#pragma once
#include "passes.cuh"

namespace sdns {

// transform lengths with a compiled kernel: 2^k (8..2048) and 3*2^k (12..3072, the 3/2-rule lengths)
#ifndef SDNS_SIZES
#define SDNS_SIZES(X) X(8) X(12) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) \
                      X(24) X(48) X(96) X(192) X(384) X(768) X(1536) X(3072)
#endif

enum Family {
    FAM_PLAIN_FWD = 0, FAM_PLAIN_BWD = 1, FAM_NS_B0 = 2, FAM_VV_B0 = 3, FAM_NS_F0 = 4, FAM_VV_F0 = 5,
    FAM_MHD_F0 = 6, FAM_Z_C2R = 7, FAM_Z_R2C = 8, FAM_Z_CROSS = 9, FAM_Z_MHD = 10, FAM_COUNT = 11
};

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }

// launch geometry of the strided kernels
template <typename T, int N, int MODE>
struct SCfg {
    static constexpr bool heavy = (MODE == S_NS_F0 || MODE == S_VV_F0);
    static constexpr int E0 = (N % 3 == 0) ? 12 : 8;
    static constexpr int E = (N / E0 > 64) ? 2 * E0 : E0;
    static constexpr int P = N / E;
    static constexpr int maxThreads = sizeof(T) == 8 ? (heavy ? 256 : 512) : (heavy ? 512 : 1024);
    static constexpr int TCfull = 128 / (2 * (int)sizeof(T));
    static constexpr int TC = cmin(TCfull, cmax(1, maxThreads / P));
    static constexpr size_t bytes1 = (size_t)N * TC * 2 * sizeof(T);
    static constexpr int NBUF = (2 * bytes1 <= 100 * 1024) ? 2 : 1;
    static constexpr size_t smem = bytes1 * NBUF;
};

// MHD epilogue: six accumulators per thread -> fewer elements per thread
template <typename T, int N>
struct MCfg {
    static constexpr int E = (N % 3 == 0) ? (sizeof(T) == 8 ? 6 : 12) : (sizeof(T) == 8 ? 4 : 8);
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
    static constexpr int E0 = (M % 3 == 0) ? 12 : 8;
    static constexpr int E = (M / E0 > 128) ? 2 * E0 : E0;
    static constexpr int P = M / E;
    static constexpr int LPC = cmax(1, 128 / P);
    static constexpr int SYNC = (P <= 32) ? 1 : 0;
    static constexpr int NBUF = 2;
    static constexpr int PADW = 128 / (2 * (int)sizeof(T));
    static constexpr int LP = M + M / PADW + 1;
    static constexpr size_t smem = (size_t)LP * LPC * NBUF * 2 * sizeof(T);
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
#undef SDNS_DECL

}  // namespace sdns
