// inst.cu -- one instantiation unit: compile with -DSDNS_FAMILY=<0..16> -DSDNS_PREC=<32|64>.
// Defines sdns_launch_<family>_f<prec>(n, args, stream): picks the kernel compiled for transform
// length n and launches it.
#include "launch.cuh"
#include <algorithm>

#ifndef SDNS_FAMILY
#error "compile with -DSDNS_FAMILY=n"
#endif
#if SDNS_PREC == 32
typedef float real_t;
#define SDNS_FN2(f) sdns_launch_##f##_f32
#else
typedef double real_t;
#define SDNS_FN2(f) sdns_launch_##f##_f64
#endif
#define SDNS_FN1(f) SDNS_FN2(f)
#define SDNS_FN SDNS_FN1(SDNS_FAMILY)

namespace sdns {

template <typename T, int N, int MODE, int DIR>
static int run_strided(const StridedArgs<T>& a, cudaStream_t st) {
    typedef SCfg<T, N, MODE> C;
    static_assert(plan_ok(N, C::E), "no radix plan");
    auto kern = strided_kernel<T, N, C::E, C::TC, DIR, MODE, C::NBUF, C::minBlocks>;
    static bool once = false;
    if (!once) { cudaError_t e = set_smem(kern, C::smem); if (e != cudaSuccess) return (int)e; once = true; }
    long long tiles = (a.ncols + C::TC - 1) / C::TC;
    const int ny = MODE == S_PLAIN ? a.nfields : 1;
    if (a.grid_cap > 0) {
        static int nsm = 0;
        if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
        const long long cap = ((long long)a.grid_cap * nsm + ny - 1) / ny;
        if (tiles > cap) tiles = cap;
    }
    {   // the passes address rows with 32-bit element offsets relative to the column base; a pass that stores into the
        // slab owners (xchunk > 0) offsets each destination's base by the row inside that destination's chunk only
        auto mag = [](long long v) { return v < 0 ? -v : v; };
        const long long orows = a.xchunk > 0 ? a.xchunk : N;
        if ((long long)N * mag(a.in_ls) >= (1LL << 31) || orows * std::max(mag(a.out_ls), mag(a.out_ls2)) >= (1LL << 31)) return -1001;
    }
    StridedArgs<T> b = a;
    b.xuniform = 0;
    if (a.xchunk > 0 && a.xchunk % C::P == 0 && a.omap.shift % C::P == 0) {
        const int sp = (a.omap.shift + a.xchunk - 1) / a.xchunk * a.xchunk;     // shift rounded up to whole chunks
        b.xuniform = 1; b.xhi_d0 = -(sp / a.xchunk); b.xhi_b0 = sp - a.omap.shift;
    }
    // cyclic axis-1 ownership: closed-form rows when the ranks divide the threads of a line, else the tables
    b.icyc = (a.itab && a.cycP > 0 && C::P % a.cycP == 0) ? 1 : 0;
    b.ocyc = (a.otab && a.cycP > 0 && C::P % a.cycP == 0) ? 1 : 0;
#ifdef SDNS_CYC_TABLES
    b.icyc = b.ocyc = 0;
#endif
    xfer_prepare(b.x, C::smem, C::P * C::TC);
    dim3 grid((unsigned)tiles + b.x.nctas, ny);
    if constexpr (MODE == S_NS_B0 || MODE == S_VV_B0 || MODE == S_NS_GRAD_B0) {
        if (a.otab) {       // uneven split of the x0 planes: the instantiation that looks every output's owner up
            auto kx = strided_kernel<T, N, C::E, C::TC, DIR, MODE, C::NBUF, C::minBlocks, true>;
            static bool oncex = false;
            if (!oncex) { cudaError_t e = set_smem(kx, C::smem); if (e != cudaSuccess) return (int)e; oncex = true; }
            SDNS_LAUNCH(kx, grid, C::P * C::TC, C::smem, st)(b);
            return (int)cudaGetLastError();
        }
    }
    SDNS_LAUNCH(kern, grid, C::P * C::TC, C::smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int N, int MODE>
static int run_b0x(const StridedArgs<T>& a, cudaStream_t st) {
    typedef BXCfg<T, N> C;
    auto kern = a.xchunk > 0 ? b0x_kernel<T, N, C::E, C::TC, MODE, true, C::minBlocks>
                             : b0x_kernel<T, N, C::E, C::TC, MODE, false, C::minBlocks>;
    static bool once = false;
    if (!once) {
        cudaError_t e = set_smem(b0x_kernel<T, N, C::E, C::TC, MODE, true, C::minBlocks>, C::smem);
        if (e == cudaSuccess) e = set_smem(b0x_kernel<T, N, C::E, C::TC, MODE, false, C::minBlocks>, C::smem);
        if (e != cudaSuccess) return (int)e;
        once = true;
    }
    const long long tiles = (a.ncols + C::TC - 1) / C::TC;
    {   // the passes address rows with 32-bit element offsets relative to the column base; a pass that stores into the
        // slab owners (xchunk > 0) offsets each destination's base by the row inside that destination's chunk only
        auto mag = [](long long v) { return v < 0 ? -v : v; };
        const long long orows = a.xchunk > 0 ? a.xchunk : N;
        if ((long long)N * mag(a.in_ls) >= (1LL << 31) || orows * std::max(mag(a.out_ls), mag(a.out_ls2)) >= (1LL << 31)) return -1001;
    }
    StridedArgs<T> b = a;
    b.xuniform = 0;
    if (a.xchunk > 0 && a.xchunk % C::P == 0 && a.omap.shift % C::P == 0) {
        const int sp = (a.omap.shift + a.xchunk - 1) / a.xchunk * a.xchunk;     // shift rounded up to whole chunks
        b.xuniform = 1; b.xhi_d0 = -(sp / a.xchunk); b.xhi_b0 = sp - a.omap.shift;
    }
    xfer_prepare(b.x, C::smem, C::threads);
    dim3 grid((unsigned)tiles + b.x.nctas);
    SDNS_LAUNCH(kern, grid, C::threads, C::smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int N, int MODE>
static int run_b0(const StridedArgs<T>& a, cudaStream_t st) {
    // one tile per CTA: a capped grid (grid_cap, an experiment of the multi-GPU pipeline) and tile counts beyond 32 bits
    // stay with strided_kernel
    if constexpr (BXCfg<T, N>::ok) {
        if (a.grid_cap <= 0 && a.ncols < (1LL << 31) && !a.otab) return run_b0x<T, N, MODE>(a, st);
    }
    return run_strided<T, N, MODE, +1>(a, st);
}

template <typename T, int N, int MODE>
static int run_f0x(const StridedArgs<T>& a, cudaStream_t st) {
    typedef FXCfg<T, N> C;
    auto kern = f0x_kernel<T, N, C::E, C::TC, MODE, C::minBlocks>;
    static bool once = false;
    if (!once) { cudaError_t e = set_smem(kern, C::smem); if (e != cudaSuccess) return (int)e; once = true; }
    StridedArgs<T> b = a;
    xfer_prepare(b.x, C::smem, C::threads);
    dim3 grid((unsigned)((a.ncols + C::TC - 1) / C::TC) + b.x.nctas);
    SDNS_LAUNCH(kern, grid, C::threads, C::smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int N, int MODE>
static int run_f0(const StridedArgs<T>& a, cudaStream_t st) {
#ifndef SDNS_NO_F0X
    if constexpr (FXCfg<T, N>::ok) return run_f0x<T, N, MODE>(a, st);
    else
#endif
    return run_strided<T, N, MODE, -1>(a, st);
}

template <typename T, int N>
static int run_mhd_f0(const StridedArgs<T>& a, cudaStream_t st) {
    typedef MCfg<T, N> C;
    static_assert(plan_ok(N, C::E), "no radix plan");
    auto kern = mhd_f0_kernel<T, N, C::E, C::TC, C::NBUF>;
    static bool once = false;
    if (!once) { cudaError_t e = set_smem(kern, C::smem); if (e != cudaSuccess) return (int)e; once = true; }
    StridedArgs<T> b = a;
    xfer_prepare(b.x, C::smem, C::P * C::TC);
    dim3 grid((unsigned)((a.ncols + C::TC - 1) / C::TC) + b.x.nctas);
    SDNS_LAUNCH(kern, grid, C::P * C::TC, C::smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int N>
static int run_nsdiv_f0(const StridedArgs<T>& a, cudaStream_t st) {
    typedef MCfg<T, N> C;
    static_assert(plan_ok(N, C::E), "no radix plan");
    auto kern = nsdiv_f0_kernel<T, N, C::E, C::TC, C::NBUF>;
    static bool once = false;
    if (!once) { cudaError_t e = set_smem(kern, C::smem); if (e != cudaSuccess) return (int)e; once = true; }
    StridedArgs<T> b = a;
    xfer_prepare(b.x, C::smem, C::P * C::TC);
    dim3 grid((unsigned)((a.ncols + C::TC - 1) / C::TC) + b.x.nctas);
    SDNS_LAUNCH(kern, grid, C::P * C::TC, C::smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int M, int QN>
static int run_zx_q(const ZArgs<T>& a, cudaStream_t st) {
    typedef ZXCfg<T, M> C;
    auto kern = zx_kernel<T, M, C::E, C::LPC, QN, C::minBlocks>;
    static int occ_sm = 0, nsm = 0;
    if (!occ_sm) {
        cudaError_t e = set_smem(kern, C::smem); if (e != cudaSuccess) return (int)e;
        int dev = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * C::LPC, C::smem);
        if (e != cudaSuccess) return (int)e;
        occ_sm = occ > 0 ? occ : 1;
    }
    const long long blocks_per_sm = (long long)((a.grid_cap > 0 && a.grid_cap < occ_sm) ? a.grid_cap : occ_sm) * nsm;
    const long long want = (a.nlines + C::LPC - 1) / C::LPC;
    ZArgs<T> b = a;
    xfer_prepare(b.x, C::smem, 32 * C::LPC);
    // the transfer CTAs take resident slots of this persistent grid
    const long long room = std::max(blocks_per_sm - b.x.nctas, (long long)nsm);
    dim3 grid((unsigned)(want < room ? want : room) + b.x.nctas);
    SDNS_LAUNCH(kern, grid, 32 * C::LPC, C::smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int M>
static int run_zx(const ZArgs<T>& a, cudaStream_t st) {
    typedef ZXCfg<T, M> C;
    if constexpr (C::ok) {
        if (a.nin_keep <= 32 * C::QN3) return run_zx_q<T, M, C::QN3>(a, st);
        return run_zx_q<T, M, C::QN2>(a, st);
    } else {
        return -1;
    }
}

template <typename T, int M, int QN>
static int run_zb_q(const ZArgs<T>& a, cudaStream_t st) {
    typedef ZBCfg<T, M> C;
    constexpr size_t smem = C::smem_q(QN);
    auto kern = zb_kernel<T, M, C::E, C::LPC, QN, C::NST, C::minBlocks_q(QN)>;
    static int occ_sm = 0, nsm = 0;
    if (!occ_sm) {
        cudaError_t e = set_smem(kern, smem); if (e != cudaSuccess) return (int)e;
        int dev = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * C::LPC, smem);
        if (e != cudaSuccess) return (int)e;
        occ_sm = occ > 0 ? occ : 1;
    }
    const long long blocks_per_sm = (long long)((a.grid_cap > 0 && a.grid_cap < occ_sm) ? a.grid_cap : occ_sm) * nsm;
    const long long want = (a.nlines + C::LPC - 1) / C::LPC;
    ZArgs<T> b = a;
    xfer_prepare(b.x, smem, 32 * C::LPC);
    const long long room = std::max(blocks_per_sm - b.x.nctas, (long long)nsm);
    dim3 grid((unsigned)(want < room ? want : room) + b.x.nctas);
    SDNS_LAUNCH(kern, grid, 32 * C::LPC, smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int M>
static int run_zb(const ZArgs<T>& a, cudaStream_t st) {
    typedef ZXCfg<T, M> X;
    if constexpr (ZBCfg<T, M>::ok) {
        // the staged rows are K2p = in_ls elements long
        if (a.in_ls <= 32 * X::QN3) return run_zb_q<T, M, X::QN3>(a, st);
        if (a.in_ls <= 32 * X::QN2) return run_zb_q<T, M, X::QN2>(a, st);
        return run_zb_q<T, M, X::QN2 + 1>(a, st);
    } else {
        return -1;
    }
}

template <typename T, int M, int QN>
static int run_zy_q(const ZArgs<T>& a, cudaStream_t st) {
    typedef ZYCfg<T, M> C;
    constexpr size_t smem = C::smem_q(QN);
    auto kern = zy_kernel<T, M, C::E, QN, C::minBlocks_q(QN)>;
    static int occ_sm = 0, nsm = 0;
    if (!occ_sm) {
        cudaError_t e = set_smem(kern, smem); if (e != cudaSuccess) return (int)e;
        int dev = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::P, smem);
        if (e != cudaSuccess) return (int)e;
        occ_sm = occ > 0 ? occ : 1;
    }
    const long long blocks = (long long)((a.grid_cap > 0 && a.grid_cap < occ_sm) ? a.grid_cap : occ_sm) * nsm;
    ZArgs<T> b = a;
    xfer_prepare(b.x, smem, C::P);
    const long long room = std::max(blocks - b.x.nctas, (long long)nsm);
    dim3 grid((unsigned)(a.nlines < room ? a.nlines : room) + b.x.nctas);
    SDNS_LAUNCH(kern, grid, C::P, smem, st)(b);
    return (int)cudaGetLastError();
}

template <typename T, int M>
static int run_zy(const ZArgs<T>& a, cudaStream_t st) {
    typedef ZYCfg<T, M> C;
    if constexpr (C::ok) {
        if (a.nin_keep <= C::P * C::QN3) return run_zy_q<T, M, C::QN3>(a, st);
        return run_zy_q<T, M, C::QN2>(a, st);
    } else {
        return -1;
    }
}

template <typename T, int M, int MODE>
static int run_z(const ZArgs<T>& a, cudaStream_t st) {
#ifdef SDNS_ZY_E4
    if constexpr (MODE == Z_CROSS && ZYCfg<T, M>::ok && ZYCfg<T, M>::E == 4) return run_zy<T, M>(a, st);
#endif
    if constexpr (MODE == Z_CROSS && ZXCfg<T, M>::ok) {
#ifdef SDNS_USE_ZB      // bulk-async staged z pass: measured, not faster than zx_kernel; opt-in
        if ((a.in_ls * (long long)(2 * sizeof(T))) % 16 == 0 && (a.in_fs * (long long)(2 * sizeof(T))) % 16 == 0 && ((uintptr_t)a.in % 16) == 0)
            return run_zb<T, M>(a, st);
#endif
#ifndef SDNS_NO_ZX
        return run_zx<T, M>(a, st);
#endif
    }
    if constexpr (MODE == Z_CROSS && ZYCfg<T, M>::ok) {
#ifndef SDNS_NO_ZY
        return run_zy<T, M>(a, st);
#endif
    }
    typedef ZCfg<T, M, MODE> C;
    static_assert(plan_ok(M, C::E), "no radix plan");
    auto kern = z_kernel<T, M, C::E, C::LPC, MODE, C::SYNC, C::NBUF, (C::minBlocks < 1 ? 1 : C::minBlocks)>;
    static bool once = false;
    if (!once) { cudaError_t e = set_smem(kern, C::smem); if (e != cudaSuccess) return (int)e; once = true; }
    ZArgs<T> b = a;
    xfer_prepare(b.x, C::smem, C::P * C::LPC);
    dim3 grid((unsigned)((a.nlines + C::LPC - 1) / C::LPC) + b.x.nctas);
    SDNS_LAUNCH(kern, grid, C::P * C::LPC, C::smem, st)(b);
    return (int)cudaGetLastError();
}

int SDNS_FN(int n, const void* args, cudaStream_t st) {
    typedef real_t T;
    switch (n) {
#if SDNS_FAMILY == 0
#define X(N) case N: return run_strided<T, N, S_PLAIN, -1>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 1
#define X(N) case N: return run_strided<T, N, S_PLAIN, +1>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 2
#define X(N) case N: return run_b0<T, N, S_NS_B0>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 3
#define X(N) case N: return run_b0<T, N, S_VV_B0>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 4
#define X(N) case N: return run_f0<T, N, S_NS_F0>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 5
#define X(N) case N: return run_f0<T, N, S_VV_F0>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 6
#define X(N) case N: return run_mhd_f0<T, N>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 7
#define X(N) case N: return run_z<T, N, Z_C2R>(*(const ZArgs<T>*)args, st);
#elif SDNS_FAMILY == 8
#define X(N) case N: return run_z<T, N, Z_R2C>(*(const ZArgs<T>*)args, st);
#elif SDNS_FAMILY == 9
#define X(N) case N: return run_z<T, N, Z_CROSS>(*(const ZArgs<T>*)args, st);
#elif SDNS_FAMILY == 10
#define X(N) case N: return run_z<T, N, Z_MHD>(*(const ZArgs<T>*)args, st);
#elif SDNS_FAMILY == 11
#define X(N) case N: return run_strided<T, N, S_NS_GRAD_B0, +1>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 12
#define X(N) case N: return run_z<T, N, Z_DOT>(*(const ZArgs<T>*)args, st);
#elif SDNS_FAMILY == 13
#define X(N) case N: return run_z<T, N, Z_UU>(*(const ZArgs<T>*)args, st);
#elif SDNS_FAMILY == 14
#define X(N) case N: return run_nsdiv_f0<T, N>(*(const StridedArgs<T>*)args, st);
#elif SDNS_FAMILY == 15
#define X(N) case N: return run_z<T, N, Z_NS2D>(*(const ZArgs<T>*)args, st);
#elif SDNS_FAMILY == 16
#define X(N) case N: return run_z<T, N, Z_BQ2D>(*(const ZArgs<T>*)args, st);
#endif
        SDNS_SIZES(X)
#if SDNS_FAMILY <= 5 || (SDNS_FAMILY >= 7 && SDNS_FAMILY <= 9)
        SDNS_SIZES_5(X)
#endif
#undef X
        default: return -1000;
    }
}

}  // namespace sdns
