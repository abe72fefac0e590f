// passes.cuh -- the axis passes of the pseudo-spectral right-hand side, sm_100a.
//
// A scalar 3-D transform is three axis passes (SURVEY.md section 8d "three-pass model").  All
// pointwise work of the reference's hot path is fused into them:
//   B0  axis-0 inverse c2c.  Load functor: 2/3 truncation or 3/2 zero padding (shenfun Tp,
//       solvers/NS.py:29-31), curl_hat = i k x u_hat (cross2, maths/cross.py:30-35,
//       optimization/cython_maths.in:62-86) or u_hat = i k x w_hat / k^2 (VV.py:52-67).
//   B1  axis-1 inverse c2c.
//   Z   axis-2: c2r of all fields, the real-space product (cross1 maths/cross.py:16-28 /
//       Elsasser products MHD.py:99-110), r2c of the products -- one kernel, the real fields
//       never touch HBM.
//   F1  axis-1 forward c2c (+ 3/2 truncation).
//   F0  axis-0 forward c2c.  Epilogue: Nyquist mask (NS.py:253-254), pressure projection +
//       viscous term (add_pressure_diffusion, NS.py:203-217, cython_solvers.in:44-80), Source
//       (NS.py:259), VV/MHD variants (VV.py:99-110, MHD.py:89-97,132-149) and the RK4 stage
//       update (maths/integrators.py:150-159, cython_integrators.in:8-52).
#pragma once
#include "fft_core.cuh"
#include "xfer.cuh"

namespace sdns {

// transform index j in [0, M)  <->  memory index along the same axis (or -1: not stored)
struct AxisMap { int nlo, nhi, shift; };
__device__ __forceinline__ int axis_mem(const AxisMap& a, int M, int j) {
    return j < a.nlo ? j : (j >= M - a.nhi ? j - a.shift : -1);
}
// branch-free form: memory index (garbage when !ok) and validity
__device__ __forceinline__ int axis_idx(const AxisMap& a, int j) { return j < a.nlo ? j : j - a.shift; }
__device__ __forceinline__ bool axis_ok(const AxisMap& a, int M, int j) { return j < a.nlo || j >= M - a.nhi; }

enum StridedMode { S_PLAIN = 0, S_NS_B0 = 1, S_VV_B0 = 2, S_NS_F0 = 3, S_VV_F0 = 4, S_MHD_F0 = 5, S_NS_GRAD_B0 = 6 };
enum ZMode { Z_C2R = 0, Z_R2C = 1, Z_CROSS = 2, Z_MHD = 3, Z_DOT = 4, Z_UU = 5, Z_NS2D = 6, Z_BQ2D = 7 };
enum OutMode { OUT_RHS = 0, OUT_STAGE = 1, OUT_CONV = 2 };   // OUT_CONV: convection term only (solver.conv)

template <typename T>
struct StridedArgs {
    typedef typename C2<T>::type V;
    const V* in; V* out;
    long long in_fs, out_fs;      // field strides (elements)
    long long in_ls, out_ls;      // stride along the transform line
    long long in_os, out_os;      // stride between column runs
    int cw;                       // columns per run (contiguous in memory)
    long long ncols;              // total columns
    int col_nlo, col_gap;         // run index -> memory run index (pruned axis-1 set on B0 input)
    AxisMap imap, omap;
    const V* tw;
    T scale;
    int nfields;
    // functor data
    const T* kx; const T* ky; const T* kz;      // scaled wavenumbers K[0],K[1],K[2] (NS.py:38-41)
    int N0, N1, N2;                             // logical (unpadded) grid
    int mask_nyquist;
    // epilogue (F0 modes)
    int out_mode;                 // OutMode
    const V* u_hat;               // stage input u0 (viscous term)   [may alias u0]
    V* rhs;                       // OUT_RHS target
    V* u0; V* u1; V* u2;          // OUT_STAGE state
    const V* source;              // optional
    V* p_hat;                     // optional (NS OUT_RHS)
    long long st_fs;              // field stride of the dense state arrays (= N0*N1*Nh)
    // element (i0, c1, c2) of u_hat sits at i0*uh_ls + c1*uh_os + c2, of the integrator's work arrays
    // (u1, u2, and u0 between stages) at i0*t_ls + c1*t_os + c2: the library keeps those k1-major so
    // that the axis-0 passes touch them with small strides; rhs / source / p_hat / final u0 use out_ls/out_os
    long long uh_ls, uh_os, t_ls, t_os;
    T nu, eta, adt, bdt;
    int rk;
    // slab decomposition (one process per GPU): the pass that precedes a global transpose stores
    // straight into the destination rank's buffer (NVLink peer memory), element (f, i, column) of
    // the output going to rank i / xchunk at line index i % xchunk.  xchunk == 0: single GPU.
    int comp;                     // S_NS_GRAD_B0: component whose gradient is formed (NS.py:138-145)
    const V* addin;               // divergence-form epilogue: spectral term added to the convection (Skewed)
    T cfac;                       // divergence-form epilogue: factor on the convection (-1 or -0.5)
    int xchunk;
    long long c1_out_off;         // added to the run index of the output base (global x0 / compact k1)
    int k1_off, k1_mul;           // global axis-1 index of local column j: k1_off + j*k1_mul (Nyquist test in the epilogues)
    // SDNS_K1_CYCLIC (sdns_api.cu fill_tables): row of W0 that holds transform index j of the axis-1 backward pass, and
    // (owner rank << 24 | local row) of output j of the axis-1 forward pass; -1: not kept.  Null: the AxisMaps apply.
    // otab also serves the uneven slab splits (N1 or M0 not divisible by the ranks): (owner, row) of F1's axis-1 outputs
    // and of B0's x0 outputs.
    const int* itab; const int* otab;
    // the same maps in closed form, used when the ranks P divide the threads per line (run_strided decides): a thread's
    // rows then belong to one rank per kept range and advance by (threads per line)/P.  cyc_first[r]: first W0 row of
    // rank r's kept modes, cyc_hi[r]: the same minus the rows its truncated gap skips, cyc_dm: transform index - mode
    // index in the high range.
    int icyc, ocyc, cycP, cyc_dm;
    int cyc_first[8], cyc_hi[8];
    int c1_off, c2_off;           // first run / first column of this launch (the multi-GPU pipeline launches a pass in chunks)
    int grid_cap;                 // > 0: launch at most this many CTAs per SM (grid-stride over the tiles)
    int xuniform;                 // slab stores: P and omap.shift divide into xchunk (destination uniform per q)
    int xhi_d0, xhi_b0;           // (rank, row in chunk) of transform block q = 0 mapped through the high kept range
    V* peer_out[8];
    int self;                     // destination that uses (out_fs, out_ls, c1_out_off); all others use the second set
    long long out_fs2, out_ls2, c1_out_off2;
    XferArgs x;                   // transfer role carried by this launch (xfer.cuh); x.nctas == 0: none
};

template <typename V> __device__ __forceinline__ V czero() { return V(); }

template <typename T, typename V> __device__ __forceinline__ V cscale(V a, T s) { return cscl(a, s); }

// i*(ka*b - kb*a) for real ka,kb, complex a,b   (one component of cross2)
template <typename T, typename V>
__device__ __forceinline__ V icross(T ka, V b, T kb, V a) {
    // (ka*b - kb*a) * i = ( -(ka*b.y - kb*a.y), ka*b.x - kb*a.x )
    V r; r.x = -(ka * b.y - kb * a.y); r.y = ka * b.x - kb * a.x; return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef SDNS_HOST_SHIM
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// L2 prefetch of one transform line (costs issue slots but no registers): used where a kernel reads
// several fields one after the other, so that only the first one pays DRAM latency.
template <typename T, int N, int E, typename V>
__device__ __forceinline__ void prefetch_line(const V* __restrict__ pin, long long ls, const AxisMap& m, int t, bool valid) {
    constexpr int P = N / E;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int j = t + q * P;
        if (valid && axis_ok(m, N, j)) prefetch_l2(pin + (long long)axis_idx(m, j) * ls);
    }
}

// Hides how a pointer was formed, so that the compiler keeps it in a register pair and addresses row r as one
// IMAD.WIDE off it instead of re-deriving a 64-bit element index per access.
template <typename Q> __device__ __forceinline__ Q* opaque(Q* p) {
#ifndef SDNS_HOST_SHIM
    unsigned long long v = reinterpret_cast<unsigned long long>(p);
    asm volatile("" : "+l"(v));
    Q* q = reinterpret_cast<Q*>(v);
    __builtin_assume(__isGlobal(q));
    return q;
#else
    return p;
#endif
}

// Row bookkeeping of one thread of a strided pass.  The thread owns the transform indices j = t + q P, q < E.
// An AxisMap keeps j < nlo at memory row j (low block) and j >= N - nhi at memory row j - shift (high block), so
// per thread there are two q-ranges, q < qlo and q >= qhi, two row-base pointers, and one uniform stride P*ls
// between consecutive q: an access costs a compare pair, a select and one 64-bit add.
template <int N, int E>
struct RowSel {
    int qlo, qhi;
    __device__ __forceinline__ RowSel(const AxisMap& m, int t, bool valid) {
        constexpr int P = N / E;
        qlo = valid ? (m.nlo - t + P - 1) / P : 0;           // q < qlo   <=>  t + q P < nlo
        qhi = valid ? (N - m.nhi - t + P - 1) / P : E;       // q >= qhi  <=>  t + q P >= N - nhi
    }
    __device__ __forceinline__ bool lo(int q) const { return q < qlo; }
    __device__ __forceinline__ bool ok(int q) const { return q < qlo || q >= qhi; }
};

// Loads of one transform line.  Row offsets are 32-bit element counts relative to the (64-bit) column base: the
// host checks N * ls < 2^31.
template <typename T, int N, int E, typename V>
__device__ __forceinline__ void load_line(V (&x)[E], const V* __restrict__ pin, long long ls, const AxisMap& m,
                                          int t, bool valid) {
    constexpr int P = N / E;
    const RowSel<N, E> rs(m, t, valid);
    const int l = (int)ls;
    const int olo = t * l, ohi = (t - m.shift) * l, step = P * l;
    const V* __restrict__ pb = opaque(pin);
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int off = (rs.lo(q) ? olo : ohi) + q * step;
        V v = czero<V>();
        if (rs.ok(q)) v = pb[off];
        x[q] = v;
    }
}

// Table-driven variants for the cyclic slab ownership of axis 1: the row of every transform index comes from a small
// table (L1-resident, shared by all columns) instead of the two-range AxisMap.
template <typename T, int N, int E, typename V>
__device__ __forceinline__ void load_line_tab(V (&x)[E], const V* __restrict__ pin, long long ls,
                                              const int* __restrict__ tab, int t, bool valid) {
    constexpr int P = N / E;
    const int l = (int)ls;
    const V* __restrict__ pb = opaque(pin);
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int r = valid ? tab[t + q * P] : -1;
        V v = czero<V>();
        if (r >= 0) v = pb[r * l];
        x[q] = v;
    }
}
// Closed-form variants (P | threads per line): the loads keep load_line's shape -- two row bases, one stride.
template <typename T, int N, int E, typename V>
__device__ __forceinline__ void load_line_cyc(V (&x)[E], const V* __restrict__ pin, long long ls, const StridedArgs<T>& a,
                                              int t, bool valid) {
    constexpr int P = N / E;
    const RowSel<N, E> rs(a.imap, t, valid);
    const int R = a.cycP, l = (int)ls;
    const int rl = t % R;
    const int th = t - a.cyc_dm;
    const int rh = ((th % R) + R) % R;
    const int olo = (a.cyc_first[rl] + t / R) * l, ohi = (a.cyc_hi[rh] + (th - rh) / R) * l, step = (P / R) * l;
    const V* __restrict__ pb = opaque(pin);
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int off = (rs.lo(q) ? olo : ohi) + q * step;
        V v = czero<V>();
        if (rs.ok(q)) v = pb[off];
        x[q] = v;
    }
}
template <typename T, int N, int E, bool SCALE, typename V>
__device__ __forceinline__ void store_line_cyc(const V (&x)[E], const StridedArgs<T>& a, int f, long long obase,
                                               long long obase2, int t, bool valid, T scale) {
    constexpr int P = N / E;
    const RowSel<N, E> rs(a.omap, t, valid);
    const int R = a.cycP;
    const long long fo = f * a.out_fs + obase, fo2 = f * a.out_fs2 + obase2;
    const int dl = t % R;                                    // owner of this thread's outputs in the low kept range
    const int th = t - a.omap.shift;
    const int dh = ((th % R) + R) % R;                       // ... and in the high one
    const bool sl = dl == a.self, sh = dh == a.self;
    const int ll = (int)(sl ? a.out_ls : a.out_ls2), lh = (int)(sh ? a.out_ls : a.out_ls2);
    V* plo = opaque(reinterpret_cast<V*>(a.peer_out[dl]) + (sl ? fo : fo2) + (long long)(t / R) * ll);
    V* phi = opaque(reinterpret_cast<V*>(a.peer_out[dh]) + (sh ? fo : fo2)) + (long long)((th - dh) / R) * lh;
    const int slo = (P / R) * ll, shi = (P / R) * lh;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        V* pq = rs.lo(q) ? plo + q * slo : phi + (long long)q * shi;
        if (rs.ok(q)) *pq = SCALE ? cscale<T>(x[q], scale) : x[q];
    }
}

template <typename T, int N, int E, bool SCALE, typename V>
__device__ __forceinline__ void store_line_tab(const V (&x)[E], const StridedArgs<T>& a, int f, long long obase,
                                               long long obase2, int t, bool valid, T scale) {
    constexpr int P = N / E;
    const long long fo = f * a.out_fs + obase, fo2 = f * a.out_fs2 + obase2;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int e = valid ? a.otab[t + q * P] : -1;
        if (e >= 0) {
            const int dest = e >> 24, row = e & 0xffffff;
            const bool self = dest == a.self;
            V* pq = reinterpret_cast<V*>(a.peer_out[dest]) + (self ? fo : fo2) + (long long)row * (self ? a.out_ls : a.out_ls2);
            *pq = SCALE ? cscale<T>(x[q], scale) : x[q];
        }
    }
}

// Stores of one transform line.  Single GPU: pout + i*ls.  Slab decomposition: element i goes to rank i / xchunk
// at line index i % xchunk.  When P and the map's shift divide into xchunk (a.xuniform, the normal case) the
// destination rank and the chunk-local block of every q are the same for all threads of the CTA, so they are
// computed once in uniform registers; otherwise per element (reciprocal multiply, exact for i < 2^22).
template <typename T, int N, int E, bool SCALE, typename V, bool LOCAL_ONLY = false>
__device__ __forceinline__ void store_line(const V (&x)[E], const StridedArgs<T>& a, int f, long long obase,
                                           long long obase2, int t, bool valid, T scale) {
    constexpr int P = N / E;
    const long long fo = f * a.out_fs + obase;
    const RowSel<N, E> rs(a.omap, t, valid);
    if (LOCAL_ONLY || a.xchunk == 0) {          // LOCAL_ONLY: the kernel variant launched on a single GPU
        V* pout = opaque(reinterpret_cast<V*>(a.out) + fo);
        const int l = (int)a.out_ls;
        const int olo = t * l, ohi = (t - a.omap.shift) * l, step = P * l;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int off = (rs.lo(q) ? olo : ohi) + q * step;
            if (rs.ok(q)) pout[off] = SCALE ? cscale<T>(x[q], scale) : x[q];
        }
    } else if (a.xuniform) {
        // destination `self` (this rank, or -1) uses the layout of the final array; every other one the second
        // layout (fs2, ls2, obase2): identical to the first for direct peer stores, the send-buffer layout when
        // the copy engines carry the exchange
        const long long fo2 = f * a.out_fs2 + obase2;
        const int l1 = (int)a.out_ls, l2 = (int)a.out_ls2;
        // (destination rank, first row inside its chunk) of block q in the low and in the high kept range: the same
        // for every thread, advanced by P rows per q
        int dl = 0, bl = 0, dh = a.xhi_d0, bh = a.xhi_b0;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const bool lo = rs.lo(q);
            int dest = lo ? dl : dh;
            dest = dest < 0 ? 0 : dest;
            const int row = t + (lo ? bl : bh);
            const bool self = dest == a.self;
            V* pq = reinterpret_cast<V*>(a.peer_out[dest]) + (self ? fo : fo2) + row * (self ? l1 : l2);
            if (rs.ok(q)) *pq = SCALE ? cscale<T>(x[q], scale) : x[q];
            bl += P; if (bl >= a.xchunk) { bl -= a.xchunk; ++dl; }
            bh += P; if (bh >= a.xchunk) { bh -= a.xchunk; ++dh; }
        }
    } else {
        const float inv = 1.0f / (float)a.xchunk;
        const long long fo2 = f * a.out_fs2 + obase2;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int j = t + q * P;
            const int i = axis_idx(a.omap, j);
            const int dest = __float2int_rz(((float)i + 0.5f) * inv);
            const int il = i - dest * a.xchunk;
            const long long off = dest == a.self ? fo + (long long)il * a.out_ls : fo2 + (long long)il * a.out_ls2;
            if (valid && axis_ok(a.omap, N, j))
                reinterpret_cast<V*>(a.peer_out[dest])[off] = SCALE ? cscale<T>(x[q], scale) : x[q];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Strided c2c pass over a tile of TC adjacent columns (TC*sizeof(V) = 128 bytes -> every
// global access of a warp covers whole 128-byte lines; thread index = column + TC*t).
// ---------------------------------------------------------------------------------------
// XTAB (B0 modes): the x0 planes are split unevenly over the ranks and every output finds (owner, local plane) in
// a.otab; a separate instantiation, so that the kernels of the even split stay exactly as they were measured.
template <typename T, int N, int E, int TC, int DIR, int MODE, int NBUF, int MINB, bool XTAB = false>
__global__ void __launch_bounds__((N / E) * TC, MINB)
strided_kernel(const StridedArgs<T> a) {
    typedef typename C2<T>::type V;
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int P = N / E;
    const int c = threadIdx.x % TC;
    const int t = threadIdx.x / TC;
    SmemLine<TC, 0> map; map.base = c;
    int phase = 0;
    constexpr int BUFSTRIDE = N * TC;
    // One tile of TC columns per CTA; launched with a capped grid (grid_cap, the multi-GPU pipeline) the CTA walks
    // the tiles with a grid stride instead, so that an NVLink-bound pass leaves SM room for the pass beside it.
    const long long ntiles = (a.ncols + TC - 1) / TC;
    for (long long tile = bx; tile < ntiles; tile += gx) {
    const long long col = tile * TC + c;
    const bool valid = col < a.ncols;
    const int c1 = valid ? (int)(col / a.cw) + a.c1_off : 0;      // c1_off / c2_off: chunked launches
    const int c2 = valid ? (int)(col % a.cw) + a.c2_off : 0;
    const int c1m = c1 < a.col_nlo ? c1 : c1 + a.col_gap;
    const long long ibase = (long long)c1m * a.in_os + c2;
    const long long obase = ((long long)c1 + a.c1_out_off) * a.out_os + c2;
    const long long obase2 = ((long long)c1 + a.c1_out_off2) * a.out_os + c2;

    if (MODE == S_PLAIN) {
        const int f = blockIdx.y;
        V x[E];
        if (a.icyc) load_line_cyc<T, N, E>(x, a.in + (f * a.in_fs + ibase), a.in_ls, a, t, valid);
        else if (a.itab) load_line_tab<T, N, E>(x, a.in + (f * a.in_fs + ibase), a.in_ls, a.itab, t, valid);
        else load_line<T, N, E>(x, a.in + (f * a.in_fs + ibase), a.in_ls, a.imap, t, valid);
        fft_line<T, N, E, DIR, 0, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
        if (a.ocyc) store_line_cyc<T, N, E, true>(x, a, f, obase, obase2, t, valid, a.scale);
        else if (a.otab) store_line_tab<T, N, E, true>(x, a, f, obase, obase2, t, valid, a.scale);
        else store_line<T, N, E, true>(x, a, f, obase, obase2, t, valid, a.scale);
    } else if (MODE == S_NS_B0 || MODE == S_VV_B0 || MODE == S_NS_GRAD_B0) {
        // in: 3 dense spectral fields.  out: 6 fields (NS: u_hat, i k x u_hat ; VV: i k x w_hat / k^2, w_hat)
        const T k1 = valid ? a.ky[c1m] : (T)0;
        const T k2 = valid ? a.kz[c2] : (T)0;
        const V* pin = a.in + ibase;
#ifndef SDNS_NO_PREFETCH
        prefetch_line<T, N, E>(pin + a.in_fs, a.in_ls, a.imap, t, valid);
        prefetch_line<T, N, E>(pin + 2 * a.in_fs, a.in_ls, a.imap, t, valid);
#endif
#pragma unroll 1
        for (int f = 0; f < 6; ++f) {
            V x[E];
            const bool direct = (MODE == S_VV_B0) ? (f >= 3) : (f < 3);
            const int g = f % 3;               // component
            const int ga = (g + 1) % 3, gb = (g + 2) % 3;
            if (direct) {
                load_line<T, N, E>(x, pin + g * a.in_fs, a.in_ls, a.imap, t, valid);
            } else if (MODE == S_NS_GRAD_B0) {
                // field 3+g = 1j*K[g]*u_hat[comp]   (standard_convection, NS.py:138-145)
                const V* pc = pin + a.comp * a.in_fs;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const int j = t + q * P;
                    const int i = axis_idx(a.imap, j);
                    const bool ok = valid && axis_ok(a.imap, N, j);
                    V b = czero<V>();
                    T k0 = (T)0;
                    if (ok) { b = pc[(long long)i * a.in_ls]; k0 = a.kx[i]; }
                    const T kg = g == 0 ? k0 : (g == 1 ? k1 : k2);
                    x[q].x = -kg * b.y; x[q].y = kg * b.x;
                }
            } else {
                // component g of i*(K x b) = i*(K[ga]*b[gb] - K[gb]*b[ga])
                const V* pa = pin + ga * a.in_fs;
                const V* pb = pin + gb * a.in_fs;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const int j = t + q * P;
                    const int i = axis_idx(a.imap, j);
                    const bool ok = valid && axis_ok(a.imap, N, j);
                    const long long off = (long long)i * a.in_ls;
                    V bb = czero<V>(), ba = czero<V>();
                    T k0 = (T)0;
                    if (ok) { bb = pb[off]; ba = pa[off]; k0 = a.kx[i]; }
                    T ka = ga == 0 ? k0 : (ga == 1 ? k1 : k2);
                    T kb = gb == 0 ? k0 : (gb == 1 ? k1 : k2);
                    if (MODE == S_VV_B0) {
                        T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;   // NS.py:42-44
                        if (ksq == (T)0) ksq = (T)1;                        // NS.py:46-48
                        ka = ka / ksq; kb = kb / ksq;                       // K_over_K2
                    }
                    x[q] = icross<T, V>(ka, bb, kb, ba);
                }
            }
            fft_line<T, N, E, DIR, 0, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
            if (XTAB) store_line_tab<T, N, E, false>(x, a, f, obase, obase2, t, valid, (T)1);       // uneven split of x0
            else store_line<T, N, E, false>(x, a, f, obase, obase2, t, valid, (T)1);
        }
    } else if (MODE == S_NS_F0 || MODE == S_VV_F0) {
        // Three forward transforms; the results of the first two are parked in thread-private
        // shared-memory slots ([field][q][thread], conflict free) so that only one field lives in
        // registers: ~100 instead of 172 registers -> two resident CTAs per SM whose load /
        // transform / epilogue phases overlap.
        constexpr int NT = P * TC;
        V* park = sm + NBUF * BUFSTRIDE;
        V x[E];
#pragma unroll 1
        for (int f = 0; f < 3; ++f) {
            load_line<T, N, E>(x, a.in + (f * a.in_fs + ibase), a.in_ls, a.imap, t, valid);
            fft_line<T, N, E, DIR, 0, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
            if (f < 2) {
#pragma unroll
                for (int q = 0; q < E; ++q) park[(f * E + q) * NT + threadIdx.x] = x[q];
            }
        }
        if (!valid) continue;
        const int i1 = c1, i2 = c2;
        const T k1 = a.ky[i1], k2 = a.kz[i2];
        const bool nyq12 = a.mask_nyquist && ((2 * (i1 * a.k1_mul + a.k1_off) == a.N1) || (2 * i2 == a.N2));
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int i0 = axis_mem(a.omap, N, t + q * P);
            if (i0 < 0) continue;
            const long long off = (long long)i0 * a.out_ls + obase;
            const T k0 = a.kx[i0];
            T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;
            V d0 = cscale<T>(park[q * NT + threadIdx.x], a.scale), d1 = cscale<T>(park[(E + q) * NT + threadIdx.x], a.scale),
              d2 = cscale<T>(x[q], a.scale);
            if (MODE == S_VV_F0) {
                // rhs = i*(K x v_hat)   (VV.py:99)
                V e0 = icross<T, V>(k1, d2, k2, d1);
                V e1 = icross<T, V>(k2, d0, k0, d2);
                V e2 = icross<T, V>(k0, d1, k1, d0);
                d0 = e0; d1 = e1; d2 = e2;
            }
            if (a.out_mode == OUT_CONV) {
                a.rhs[off] = d0; a.rhs[a.st_fs + off] = d1; a.rhs[2 * a.st_fs + off] = d2;
                continue;
            }
            if (nyq12 || (a.mask_nyquist && 2 * i0 == a.N0)) { d0 = czero<V>(); d1 = czero<V>(); d2 = czero<V>(); }
            const long long offu = (long long)i0 * a.uh_ls + (long long)c1 * a.uh_os + c2;
            const long long offt = (long long)i0 * a.t_ls + (long long)c1 * a.t_os + c2;
            const V w0 = a.u_hat[offu], w1 = a.u_hat[a.st_fs + offu], w2 = a.u_hat[2 * a.st_fs + offu];
            const T z = a.nu * ksq;
            if (MODE == S_NS_F0) {
                const T ks = ksq == (T)0 ? (T)1 : ksq;
                const T q0 = k0 / ks, q1 = k1 / ks, q2 = k2 / ks;       // K_over_K2 (NS.py:46-48)
                V p;
                p.x = d0.x * q0 + d1.x * q1; p.x += d2.x * q2;
                p.y = d0.y * q0 + d1.y * q1; p.y += d2.y * q2;
                if (a.p_hat) a.p_hat[off] = p;
                d0.x -= p.x * k0; d0.y -= p.y * k0;
                d1.x -= p.x * k1; d1.y -= p.y * k1;
                d2.x -= p.x * k2; d2.y -= p.y * k2;
            }
            d0.x -= z * w0.x; d0.y -= z * w0.y;
            d1.x -= z * w1.x; d1.y -= z * w1.y;
            d2.x -= z * w2.x; d2.y -= z * w2.y;
            if (a.source) {
                d0 = cadd(d0, a.source[off]); d1 = cadd(d1, a.source[a.st_fs + off]);
                d2 = cadd(d2, a.source[2 * a.st_fs + off]);
            }
            V dd[3] = {d0, d1, d2};
            V ww[3] = {w0, w1, w2};
            if (a.out_mode == OUT_RHS) {
#pragma unroll
                for (int f = 0; f < 3; ++f) a.rhs[f * a.st_fs + off] = dd[f];
            } else {
                // RK4 stage rk (integrators.py:150-159): u1 = u2 = u0 at rk 0;
                // u0 = u1 + b*dt*rhs (rk<3); u2 += a*dt*rhs; u0 = u2 after rk 3.
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    const long long o = f * a.st_fs + offt;
                    V b1, b2;
                    if (a.rk == 0) { b1 = ww[f]; b2 = ww[f]; a.u1[o] = b1; }
                    else { b2 = a.u2[o]; if (a.rk < 3) b1 = a.u1[o]; }
                    b2.x += a.adt * dd[f].x; b2.y += a.adt * dd[f].y;
                    if (a.rk < 3) {
                        a.u2[o] = b2;
                        V n; n.x = b1.x + a.bdt * dd[f].x; n.y = b1.y + a.bdt * dd[f].y;
                        a.u0[o] = n;
                    } else {
                        a.u0[f * a.st_fs + off] = b2;                  // final u0: reference layout
                    }
                }
            }
        }
    }
    }   // tile loop
}

// ---------------------------------------------------------------------------------------
// F0 for NS / VV, field-parallel: the CTA has three thread groups, group g transforms field g, so a
// tile has ONE load phase and ONE transform phase (the strided_kernel F0 branch runs three of each
// back to back and is latency-bound).  After its last exchange a group parks its result in its own
// exchange buffer ([q][thread]); the epilogue points are then split across the groups.
// ---------------------------------------------------------------------------------------
template <typename T, int N, int E, int TC, int MODE, int MINB>
__global__ void __launch_bounds__(3 * (N / E) * TC, MINB)
f0x_kernel(const StridedArgs<T> a) {
    typedef typename C2<T>::type V;
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int P = N / E;
    constexpr int NT = P * TC;
    const int g = threadIdx.x / NT;                 // field handled by this thread group
    const int tid = threadIdx.x - g * NT;
    const int c = tid % TC;
    const int t = tid / TC;
    const long long col = (long long)bx * TC + c;
    const bool valid = col < a.ncols;
    const int c1 = valid ? (int)(col / a.cw) + a.c1_off : 0;      // c1_off / c2_off: chunked launches
    const int c2 = valid ? (int)(col % a.cw) + a.c2_off : 0;
    const long long ibase = (long long)c1 * a.in_os + c2;
    const long long obase = (long long)c1 * a.out_os + c2;
    SmemLine<TC, 0> map; map.base = c;
    int phase = 0;
    V* ex = sm + g * (N * TC);
#ifdef SDNS_F0_PREFETCH
    // The epilogue moves four times the bytes of the transform input (the RK4 stage update touches u_hat, u1, u2 and
    // u0).  Ask L2 for this thread's epilogue points now: the state streams in from HBM underneath the W3 loads and
    // the transform (which has no memory traffic of its own) instead of after them, at no register cost.
    if (valid && a.out_mode == OUT_STAGE) {
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int j0 = t + q * P;
            if (q % 3 == g % 3 && axis_ok(a.omap, N, j0)) {
                const int i0 = axis_idx(a.omap, j0);
                const long long offu = (long long)i0 * a.uh_ls + (long long)c1 * a.uh_os + c2;
                const long long offt = (long long)i0 * a.t_ls + (long long)c1 * a.t_os + c2;
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    prefetch_l2(a.u_hat + f * a.st_fs + offu);
                    if (a.rk > 0) prefetch_l2(a.u2 + f * a.st_fs + offt);
                    if (a.rk > 0 && a.rk < 3) prefetch_l2(a.u1 + f * a.st_fs + offt);
                }
            }
        }
    }
#endif
    {
        V x[E];
        load_line<T, N, E>(x, a.in + (g * a.in_fs + ibase), a.in_ls, a.imap, t, valid);
        fft_line<T, N, E, -1, 0, 1>(x, t, a.tw, ex, map, 0, phase);
        __syncthreads();                              // every gather from the exchange buffers is done
#pragma unroll
        for (int q = 0; q < E; ++q) ex[q * NT + tid] = x[q];
    }
    __syncthreads();
    if (!valid) return;
    const V* pk0 = sm;
    const V* pk1 = sm + N * TC;
    const V* pk2 = sm + 2 * N * TC;
    const int i1 = c1, i2 = c2;
    const T k1 = a.ky[i1], k2 = a.kz[i2];
    const bool nyq12 = a.mask_nyquist && ((2 * (i1 * a.k1_mul + a.k1_off) == a.N1) || (2 * i2 == a.N2));
#ifdef SDNS_F0_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int q = g; q < E; q += 3) {
        const int j0 = t + q * P;
        if (!axis_ok(a.omap, N, j0)) continue;
        const int i0 = axis_idx(a.omap, j0);
        const long long off = (long long)i0 * a.out_ls + obase;
        const T k0 = a.kx[i0];
        T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;
        V d0 = cscale<T>(pk0[q * NT + tid], a.scale), d1 = cscale<T>(pk1[q * NT + tid], a.scale),
          d2 = cscale<T>(pk2[q * NT + tid], a.scale);
        if (MODE == S_VV_F0) {
            // rhs = i*(K x v_hat)   (VV.py:99)
            V e0 = icross<T, V>(k1, d2, k2, d1);
            V e1 = icross<T, V>(k2, d0, k0, d2);
            V e2 = icross<T, V>(k0, d1, k1, d0);
            d0 = e0; d1 = e1; d2 = e2;
        }
        if (a.out_mode == OUT_CONV) {
            a.rhs[off] = d0; a.rhs[a.st_fs + off] = d1; a.rhs[2 * a.st_fs + off] = d2;
            continue;
        }
        if (nyq12 || (a.mask_nyquist && 2 * i0 == a.N0)) { d0 = czero<V>(); d1 = czero<V>(); d2 = czero<V>(); }
        const long long offu = (long long)i0 * a.uh_ls + (long long)c1 * a.uh_os + c2;
        const long long offt = (long long)i0 * a.t_ls + (long long)c1 * a.t_os + c2;
        const V w0 = a.u_hat[offu], w1 = a.u_hat[a.st_fs + offu], w2 = a.u_hat[2 * a.st_fs + offu];
        const T z = a.nu * ksq;
        if (MODE == S_NS_F0) {
            const T ks = ksq == (T)0 ? (T)1 : ksq;
            const T q0 = k0 / ks, q1 = k1 / ks, q2 = k2 / ks;       // K_over_K2 (NS.py:46-48)
            V p;
            p.x = d0.x * q0 + d1.x * q1; p.x += d2.x * q2;
            p.y = d0.y * q0 + d1.y * q1; p.y += d2.y * q2;
            if (a.p_hat) a.p_hat[off] = p;
            d0.x -= p.x * k0; d0.y -= p.y * k0;
            d1.x -= p.x * k1; d1.y -= p.y * k1;
            d2.x -= p.x * k2; d2.y -= p.y * k2;
        }
        d0.x -= z * w0.x; d0.y -= z * w0.y;
        d1.x -= z * w1.x; d1.y -= z * w1.y;
        d2.x -= z * w2.x; d2.y -= z * w2.y;
        if (a.source) {
            d0 = cadd(d0, a.source[off]); d1 = cadd(d1, a.source[a.st_fs + off]);
            d2 = cadd(d2, a.source[2 * a.st_fs + off]);
        }
        V dd[3] = {d0, d1, d2};
        V ww[3] = {w0, w1, w2};
        if (a.out_mode == OUT_RHS) {
#pragma unroll
            for (int f = 0; f < 3; ++f) a.rhs[f * a.st_fs + off] = dd[f];
        } else {
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const long long o = f * a.st_fs + offt;
                V b1, b2;
                if (a.rk == 0) { b1 = ww[f]; b2 = ww[f]; a.u1[o] = b1; }
                else { b2 = a.u2[o]; if (a.rk < 3) b1 = a.u1[o]; }
                b2.x += a.adt * dd[f].x; b2.y += a.adt * dd[f].y;
                if (a.rk < 3) {
                    a.u2[o] = b2;
                    V n; n.x = b1.x + a.bdt * dd[f].x; n.y = b1.y + a.bdt * dd[f].y;
                    a.u0[o] = n;
                } else {
                    a.u0[f * a.st_fs + off] = b2;                  // final u0: reference layout
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// B0 for NS / VV, field-parallel: three thread groups, group g loads state field g ONCE (one load phase per tile,
// all three fields in flight together), publishes it in thread-private shared-memory slots, transforms and stores
// it as the direct field, then forms component g of i k x (.) from the other two groups' slots, transforms and
// stores that.  strided_kernel's B0 branch runs six load -> transform -> store phases back to back per CTA, the
// three curl components re-reading their inputs; here a tile has one load phase and two transform phases.
// Shared memory: three exchange buffers + three slot arrays.  XCH: the output feeds a slab transpose (stores into
// peer GPUs / send slots); the single-GPU instance leaves that addressing out of its register budget.
// ---------------------------------------------------------------------------------------
template <typename T, int N, int E, int TC, int MODE, bool XCH, int MINB>
__global__ void __launch_bounds__(3 * (N / E) * TC, MINB)
b0x_kernel(const StridedArgs<T> a) {
    typedef typename C2<T>::type V;
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int P = N / E;
    constexpr int NT = P * TC;
    const int g = threadIdx.x / NT;                 // state field loaded by this thread group
    const int tid = threadIdx.x - g * NT;
    const int c = tid % TC;
    const int t = tid / TC;
    const int ga = (g + 1) % 3, gb = (g + 2) % 3;
    SmemLine<TC, 0> map; map.base = c;
    int phase = 0;
    V* ex = sm + g * (N * TC);                      // this group's exchange buffer
    V* raw = sm + 3 * (N * TC);                     // [field][q][thread of the group]
    const int fdir = (MODE == S_VV_B0) ? 3 + g : g; // NS: (u_hat, i k x u_hat); VV: (i k x w_hat / k^2, w_hat)
    const int fcrs = (MODE == S_VV_B0) ? g : 3 + g;
    // One tile per CTA (a grid-stride loop's state would stay live across the transforms), 32-bit tile arithmetic
    // (the host checks ncols < 2^31), and everything a store needs is re-derived from (c1, c2) right before it:
    // little more than the line itself is live across the transforms, which is what lets 768 threads share an SM.
    {
        const unsigned int ncols = (unsigned int)a.ncols;
        const unsigned int col = (unsigned int)bx * TC + c;
        const bool valid = col < ncols;
        const unsigned int run = valid ? col / (unsigned int)a.cw : 0u;
        const int c1 = (int)run + a.c1_off;
        const int c2 = (valid ? (int)(col - run * (unsigned int)a.cw) : 0) + a.c2_off;
        V x[E];
        {
            const int c1m = c1 < a.col_nlo ? c1 : c1 + a.col_gap;
            load_line<T, N, E>(x, a.in + (g * a.in_fs + (long long)c1m * a.in_os + c2), a.in_ls, a.imap, t, valid);
        }
#pragma unroll
        for (int q = 0; q < E; ++q) raw[(g * E + q) * NT + tid] = x[q];
        fft_line<T, N, E, +1, 0, 1>(x, t, a.tw, ex, map, 0, phase);
        store_line<T, N, E, false, V, !XCH>(x, a, fdir, ((long long)c1 + a.c1_out_off) * a.out_os + c2,
                                   ((long long)c1 + a.c1_out_off2) * a.out_os + c2, t, valid, (T)1);
        __syncthreads();                            // every group's slots are written
        {
            const int c1m = c1 < a.col_nlo ? c1 : c1 + a.col_gap;
            const T k1 = valid ? a.ky[c1m] : (T)0;
            const T k2 = valid ? a.kz[c2] : (T)0;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const int j = t + q * P;
                const bool ok = valid && axis_ok(a.imap, N, j);
                const V bb = raw[(gb * E + q) * NT + tid], ba = raw[(ga * E + q) * NT + tid];   // zero where not kept
                const T k0 = ok ? a.kx[axis_idx(a.imap, j)] : (T)0;
                T ka = ga == 0 ? k0 : (ga == 1 ? k1 : k2);
                T kb = gb == 0 ? k0 : (gb == 1 ? k1 : k2);
                if (MODE == S_VV_B0) {
                    T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;   // NS.py:42-44
                    if (ksq == (T)0) ksq = (T)1;                        // NS.py:46-48
                    ka = ka / ksq; kb = kb / ksq;                       // K_over_K2
                }
                x[q] = icross<T, V>(ka, bb, kb, ba);
            }
        }
        fft_line<T, N, E, +1, 0, 1>(x, t, a.tw, ex, map, 0, phase);
        store_line<T, N, E, false, V, !XCH>(x, a, fcrs, ((long long)c1 + a.c1_out_off) * a.out_os + c2,
                                   ((long long)c1 + a.c1_out_off2) * a.out_os + c2, t, valid, (T)1);
    }
}

// ---------------------------------------------------------------------------------------
// MHD F0: 9 Elsasser product fields ZZ[i][j] -> 6 rhs components (MHD.py:89-97), Nyquist mask,
// pressure projection on the first three, -nu k^2 u, -eta k^2 b (MHD.py:132-149), RK4 stage.
// Register budget: six accumulators of EH = E elements; the transforms stream through one
// work array, so E is chosen smaller than for the NS epilogue.
// ---------------------------------------------------------------------------------------
template <typename T, int N, int E, int TC, int NBUF>
__global__ void __launch_bounds__((N / E) * TC)
mhd_f0_kernel(const StridedArgs<T> a) {
    typedef typename C2<T>::type V;
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int P = N / E;
    const int c = threadIdx.x % TC;
    const int t = threadIdx.x / TC;
    const long long col = (long long)bx * TC + c;
    const bool valid = col < a.ncols;
    const int c1 = valid ? (int)(col / a.cw) + a.c1_off : 0;      // c1_off / c2_off: chunked launches
    const int c2 = valid ? (int)(col % a.cw) + a.c2_off : 0;
    const long long ibase = (long long)c1 * a.in_os + c2;
    const long long obase = (long long)c1 * a.out_os + c2;
    SmemLine<TC, 0> map; map.base = c;
    int phase = 0;
    constexpr int BUFSTRIDE = N * TC;
    const T k1 = valid ? a.ky[c1] : (T)0, k2 = valid ? a.kz[c2] : (T)0;

    V acc[6][E];
#pragma unroll
    for (int f = 0; f < 6; ++f)
#pragma unroll
        for (int q = 0; q < E; ++q) acc[f][q] = czero<V>();

    // R_ij = FFT0(ZZ[i][j]); S_a = sum_b K_b (R_ab + R_ba) ; D_a = sum_b K_b (R_ba - R_ab)
    // R_ij contributes  K_j*R_ij to S_i,  K_i*R_ij to S_j,  K_i*R_ij to D_j,  -K_j*R_ij to D_i
    // The nine fields stream through one work array.  fp64: the NEXT field's line is loaded into a second register set (E
    // is small here) before the current one is transformed, so that its DRAM latency runs under the transform and the
    // accumulation instead of in front of them (one CTA per SM: nothing else would hide it).  Measured on B200
    // (profiles/r2/final_1gpu/passbench_mhd_pipe.txt): fp64 256^3 1776 -> 1683 us, 512^3 14.9 -> 14.7 ms; fp32 (E = 8, twice
    // the registers per set, already spilling) 1494 -> 1730 us, so fp32 keeps the serial order.
#ifndef SDNS_NO_MHD_F0_PIPE
    constexpr bool PIPE = sizeof(T) == 8;
#else
    constexpr bool PIPE = false;
#endif
    V xn[PIPE ? E : 1];
    if constexpr (PIPE) load_line<T, N, E>(xn, a.in + ibase, a.in_ls, a.imap, t, valid);
#pragma unroll
    for (int ij = 0; ij < 9; ++ij) {
        const int i = ij / 3, j = ij % 3;
        V x[E];
        if constexpr (PIPE) {
#pragma unroll
            for (int q = 0; q < E; ++q) x[q] = xn[q];
            if (ij < 8) load_line<T, N, E>(xn, a.in + ((ij + 1) * a.in_fs + ibase), a.in_ls, a.imap, t, valid);
        } else {
            load_line<T, N, E>(x, a.in + (ij * a.in_fs + ibase), a.in_ls, a.imap, t, valid);
        }
        fft_line<T, N, E, -1, 0, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int i0 = axis_mem(a.omap, N, t + q * P);
            const T k0 = i0 >= 0 ? a.kx[i0] : (T)0;
            const T kk[3] = {k0, k1, k2};
            const T ki = kk[i], kj = kk[j];
            acc[i][q].x += kj * x[q].x;     acc[i][q].y += kj * x[q].y;
            acc[j][q].x += ki * x[q].x;     acc[j][q].y += ki * x[q].y;
            acc[3 + j][q].x += ki * x[q].x; acc[3 + j][q].y += ki * x[q].y;
            acc[3 + i][q].x -= kj * x[q].x; acc[3 + i][q].y -= kj * x[q].y;
        }
    }
    if (!valid) return;
    const int i1 = c1, i2 = c2;
    const bool nyq12 = a.mask_nyquist && ((2 * (i1 * a.k1_mul + a.k1_off) == a.N1) || (2 * i2 == a.N2));
    const T hs = (T)0.5 * a.scale;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int i0 = axis_mem(a.omap, N, t + q * P);
        if (i0 < 0) continue;
        const long long off = (long long)i0 * a.out_ls + obase;
        const T k0 = a.kx[i0];
        T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;
        V d[6];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            // rhs[:3] = -i/2 * S ; rhs[3:] = +i/2 * D
            d[f].x = hs * acc[f][q].y;      d[f].y = -hs * acc[f][q].x;
            d[3 + f].x = -hs * acc[3 + f][q].y; d[3 + f].y = hs * acc[3 + f][q].x;
        }
        if (a.out_mode == OUT_CONV) {
#pragma unroll
            for (int f = 0; f < 6; ++f) a.rhs[f * a.st_fs + off] = d[f];
            continue;
        }
        if (nyq12 || (a.mask_nyquist && 2 * i0 == a.N0)) {
#pragma unroll
            for (int f = 0; f < 6; ++f) d[f] = czero<V>();
        }
        const T ks = ksq == (T)0 ? (T)1 : ksq;
        const T q0 = k0 / ks, q1 = k1 / ks, q2 = k2 / ks;
        V p;
        p.x = d[0].x * q0 + d[1].x * q1; p.x += d[2].x * q2;
        p.y = d[0].y * q0 + d[1].y * q1; p.y += d[2].y * q2;
        if (a.p_hat) a.p_hat[off] = p;
        d[0].x -= p.x * k0; d[0].y -= p.y * k0;
        d[1].x -= p.x * k1; d[1].y -= p.y * k1;
        d[2].x -= p.x * k2; d[2].y -= p.y * k2;
        const T zu = a.nu * ksq, zb = a.eta * ksq;
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            const long long o = f * a.st_fs + off;
            const long long ot = f * a.st_fs + (long long)i0 * a.t_ls + (long long)c1 * a.t_os + c2;
            const V w = a.u_hat[f * a.st_fs + (long long)i0 * a.uh_ls + (long long)c1 * a.uh_os + c2];
            const T z = f < 3 ? zu : zb;
            V dd = d[f];
            dd.x -= z * w.x; dd.y -= z * w.y;
            if (a.source) dd = cadd(dd, a.source[o]);
            if (a.out_mode == OUT_RHS) {
                a.rhs[o] = dd;
            } else {
                V b1, b2;
                if (a.rk == 0) { b1 = w; b2 = w; a.u1[ot] = b1; }
                else { b2 = a.u2[ot]; if (a.rk < 3) b1 = a.u1[ot]; }
                b2.x += a.adt * dd.x; b2.y += a.adt * dd.y;
                if (a.rk < 3) {
                    a.u2[ot] = b2;
                    V n; n.x = b1.x + a.bdt * dd.x; n.y = b1.y + a.bdt * dd.y;
                    a.u0[ot] = n;
                } else {
                    a.u0[o] = b2;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// NS divergence-form F0 (NS.py:147-162,176-189): six product fields UU_ij (i<=j, order 00 01 02 11 12 22)
// -> conv_i = 1j * sum_j K_j UU_ij (+ addin_i), rhs_i = cfac * conv_i, then the usual Nyquist mask,
// pressure projection, viscous term, Source and stage update.
// ---------------------------------------------------------------------------------------
template <typename T, int N, int E, int TC, int NBUF>
__global__ void __launch_bounds__((N / E) * TC)
nsdiv_f0_kernel(const StridedArgs<T> a) {
    typedef typename C2<T>::type V;
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int P = N / E;
    const int c = threadIdx.x % TC;
    const int t = threadIdx.x / TC;
    const long long col = (long long)bx * TC + c;
    const bool valid = col < a.ncols;
    const int c1 = valid ? (int)(col / a.cw) + a.c1_off : 0;      // c1_off / c2_off: chunked launches
    const int c2 = valid ? (int)(col % a.cw) + a.c2_off : 0;
    const long long ibase = (long long)c1 * a.in_os + c2;
    const long long obase = (long long)c1 * a.out_os + c2;
    SmemLine<TC, 0> map; map.base = c;
    int phase = 0;
    constexpr int BUFSTRIDE = N * TC;
    const T k1 = valid ? a.ky[c1] : (T)0, k2 = valid ? a.kz[c2] : (T)0;
    V acc[3][E];
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
        for (int q = 0; q < E; ++q) acc[f][q] = czero<V>();
#pragma unroll
    for (int ij = 0; ij < 6; ++ij) {
        const int i = ij < 3 ? 0 : (ij < 5 ? 1 : 2);
        const int j = ij < 3 ? ij : (ij < 5 ? ij - 2 : 2);
        V x[E];
        load_line<T, N, E>(x, a.in + (ij * a.in_fs + ibase), a.in_ls, a.imap, t, valid);
        fft_line<T, N, E, -1, 0, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int j0 = t + q * P;
            const T k0 = axis_ok(a.omap, N, j0) ? a.kx[axis_idx(a.omap, j0)] : (T)0;
            const T ki = i == 0 ? k0 : (i == 1 ? k1 : k2);
            const T kj = j == 0 ? k0 : (j == 1 ? k1 : k2);
            acc[i][q].x += kj * x[q].x; acc[i][q].y += kj * x[q].y;
            if (i != j) { acc[j][q].x += ki * x[q].x; acc[j][q].y += ki * x[q].y; }
        }
    }
    if (!valid) return;
    const int i1 = c1, i2 = c2;
    const bool nyq12 = a.mask_nyquist && ((2 * (i1 * a.k1_mul + a.k1_off) == a.N1) || (2 * i2 == a.N2));
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int j0 = t + q * P;
        if (!axis_ok(a.omap, N, j0)) continue;
        const int i0 = axis_idx(a.omap, j0);
        const long long off = (long long)i0 * a.out_ls + obase;
        const T k0 = a.kx[i0];
        T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;
        V d[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            V cv; cv.x = -a.scale * acc[f][q].y; cv.y = a.scale * acc[f][q].x;       // 1j * acc
            if (a.addin) cv = cadd(cv, a.addin[f * a.st_fs + off]);
            d[f].x = a.cfac * cv.x; d[f].y = a.cfac * cv.y;
        }
        if (a.out_mode == OUT_CONV) {
#pragma unroll
            for (int f = 0; f < 3; ++f) a.rhs[f * a.st_fs + off] = d[f];
            continue;
        }
        if (nyq12 || (a.mask_nyquist && 2 * i0 == a.N0)) { d[0] = czero<V>(); d[1] = czero<V>(); d[2] = czero<V>(); }
        const T ks = ksq == (T)0 ? (T)1 : ksq;
        const T q0 = k0 / ks, q1 = k1 / ks, q2 = k2 / ks;
        V p;
        p.x = d[0].x * q0 + d[1].x * q1; p.x += d[2].x * q2;
        p.y = d[0].y * q0 + d[1].y * q1; p.y += d[2].y * q2;
        if (a.p_hat) a.p_hat[off] = p;
        const T kk[3] = {k0, k1, k2};
        const T z = a.nu * ksq;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const long long o = f * a.st_fs + off;
            const long long ot = f * a.st_fs + (long long)i0 * a.t_ls + (long long)c1 * a.t_os + c2;
            const V w = a.u_hat[f * a.st_fs + (long long)i0 * a.uh_ls + (long long)c1 * a.uh_os + c2];
            V dd = d[f];
            dd.x -= p.x * kk[f] + z * w.x; dd.y -= p.y * kk[f] + z * w.y;
            if (a.source) dd = cadd(dd, a.source[o]);
            if (a.out_mode == OUT_RHS) {
                a.rhs[o] = dd;
            } else {
                V b1, b2;
                if (a.rk == 0) { b1 = w; b2 = w; a.u1[ot] = b1; }
                else { b2 = a.u2[ot]; if (a.rk < 3) b1 = a.u1[ot]; }
                b2.x += a.adt * dd.x; b2.y += a.adt * dd.y;
                if (a.rk < 3) {
                    a.u2[ot] = b2;
                    V n; n.x = b1.x + a.bdt * dd.x; n.y = b1.y + a.bdt * dd.y;
                    a.u0[ot] = n;
                } else {
                    a.u0[o] = b2;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Contiguous-axis pass.  Lines of one CTA: LPC; thread index = t + P*line.
// Two real lines ride on one complex transform of length M (a + i b).
// ---------------------------------------------------------------------------------------
template <typename T>
struct ZArgs {
    typedef typename C2<T>::type V;
    const void* in; void* out;
    long long in_fs, out_fs;      // field strides (elements of the respective type)
    long long in_ls, out_ls;      // line pitch
    long long nlines;
    int nin_keep;                 // spectral modes present on input (c2r side)
    int nout_keep;                // spectral modes stored on output (r2c side)
    int nf;                       // fields for Z_C2R / Z_R2C
    const V* tw;
    T scale;                      // applied on the r2c side (and Z_C2R output)
    int grid_cap;                 // persistent kernels: resident CTAs per SM to use (0 = all), leaves room for a concurrent pass
    XferArgs x;                   // transfer role carried by this launch (xfer.cuh); x.nctas == 0: none
};


template <typename T, int M, int E, typename V>
__device__ __forceinline__ void load_pair(V (&x)[E], const V* __restrict__ A, const V* __restrict__ B,
                                          int t, int nkeep) {
    constexpr int P = M / E;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int k = t + q * P;
        const int kk = (2 * k <= M) ? k : M - k;
        V va = czero<V>(), vb = czero<V>();
        if (kk < nkeep) { va = A[kk]; if (B) vb = B[kk]; }
        if (2 * k > M) { va.y = -va.y; vb.y = -vb.y; }
        if (k == 0 || 2 * k == M) { va.y = 0; vb.y = 0; }     // c2r ignores Im of DC / Nyquist
        x[q].x = va.x - vb.y; x[q].y = va.y + vb.x;
    }
}

// x = FFT(c + i d).  Writes C[k], D[k], k < nkeep (<= M/2+1) scaled by s.  D may be null.
// The mirror exchange is one more step of the ping-pong sequence used by fft_line.
template <typename T, int M, int E, int SYNC, int NBUF, typename V, typename SM>
__device__ __forceinline__ void unpack_store_pair(const V (&x)[E], V* __restrict__ C, V* __restrict__ D,
                                                  int t, int nkeep, T s, V* sm, const SM& map,
                                                  int bufstride, int& phase) {
    constexpr int P = M / E;
    V* b = sm + (NBUF == 2 ? phase * bufstride : 0);
    if (NBUF == 1) line_sync<SYNC>();
#pragma unroll
    for (int q = 0; q < E; ++q) b[map(t + q * P)] = x[q];
    line_sync<SYNC>();
    if (NBUF == 2) phase ^= 1;
    const T h = (T)0.5 * s;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const int k = t + q * P;
        if (k < nkeep) {
            const V zm = b[map(k == 0 ? 0 : M - k)];
            V c, d;
            c.x = h * (x[q].x + zm.x); c.y = h * (x[q].y - zm.y);
            // d = -i/2 * (z - conj(zm))
            d.x = h * (x[q].y + zm.y); d.y = -h * (x[q].x - zm.x);
            C[k] = c;
            if (D) D[k] = d;
        }
    }
}

template <typename T, int M, int E, int LPC, int MODE, int SYNC, int NBUF, int MINB>
__global__ void __launch_bounds__((M / E) * LPC, MINB)
z_kernel(const ZArgs<T> a) {
    typedef typename C2<T>::type V;
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int P = M / E;
    constexpr int PADW = 128 / (int)sizeof(V);
    constexpr int LP = M + M / PADW + 1;            // padded line length in shared memory
    constexpr int BUFSTRIDE = LP * LPC;
    const int t = threadIdx.x % P;
    const int ln = threadIdx.x / P;
    long long line = (long long)bx * LPC + ln;
    const bool valid = line < a.nlines;
    if (!valid) line = a.nlines - 1;               // keep all threads alive for the barriers
    SmemLine<1, PADW> map; map.base = ln * LP;
    int phase = 0;

    if (MODE == Z_C2R) {
        // nf complex lines -> nf real lines (plain backward, diagnostics path NS.py:86-99)
        const V* in = reinterpret_cast<const V*>(a.in);
        T* out = reinterpret_cast<T*>(a.out);
        for (int f = 0; f < a.nf; f += 2) {
            const bool two = f + 1 < a.nf;
            V x[E];
            load_pair<T, M, E>(x, in + f * a.in_fs + line * a.in_ls,
                               two ? in + (f + 1) * a.in_fs + line * a.in_ls : nullptr, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
            if (valid) {
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    out[f * a.out_fs + line * a.out_ls + t + q * P] = x[q].x * a.scale;
                    if (two) out[(f + 1) * a.out_fs + line * a.out_ls + t + q * P] = x[q].y * a.scale;
                }
            }
        }
    } else if (MODE == Z_R2C) {
        const T* in = reinterpret_cast<const T*>(a.in);
        V* out = reinterpret_cast<V*>(a.out);
        for (int f = 0; f < a.nf; f += 2) {
            const bool two = f + 1 < a.nf;
            V x[E];
#pragma unroll
            for (int q = 0; q < E; ++q) {
                x[q].x = in[f * a.in_fs + line * a.in_ls + t + q * P];
                x[q].y = two ? in[(f + 1) * a.in_fs + line * a.in_ls + t + q * P] : (T)0;
            }
            fft_line<T, M, E, -1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
            V* C = out + f * a.out_fs + line * a.out_ls;
            V* D = two ? out + (f + 1) * a.out_fs + line * a.out_ls : nullptr;
            unpack_store_pair<T, M, E, SYNC, NBUF>(x, C, D, t, valid ? a.nout_keep : 0, a.scale, sm, map, BUFSTRIDE, phase);
        }
    } else {
        // fused: 6 spectral lines -> real space -> products -> spectral lines
        const V* in = reinterpret_cast<const V*>(a.in);
        V* out = reinterpret_cast<V*>(a.out);
        const int nk = valid ? a.nout_keep : 0;
        if (MODE == Z_CROSS) {
            // Pairs (a0,a1), (a2,b0), (b1,b2): the first two go to thread-private shared-memory
            // slots after their inverse transform, so one work array lives in registers.
            constexpr int NT = P * LPC;
            V* park = sm + NBUF * BUFSTRIDE;
            V x[E];
#pragma unroll 1
            for (int pr = 0; pr < 3; ++pr) {
                load_pair<T, M, E>(x, in + (2 * pr) * a.in_fs + line * a.in_ls,
                                   in + (2 * pr + 1) * a.in_fs + line * a.in_ls, t, a.nin_keep);
                fft_line<T, M, E, +1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
                if (pr < 2) {
#pragma unroll
                    for (int q = 0; q < E; ++q) park[(pr * E + q) * NT + threadIdx.x] = x[q];
                }
            }
            T* park_r = reinterpret_cast<T*>(park);
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const V p01 = park[q * NT + threadIdx.x], p23 = park[(E + q) * NT + threadIdx.x];
                const T a0 = p01.x, a1 = p01.y, a2 = p23.x;
                const T b0 = p23.y, b1 = x[q].x, b2 = x[q].y;
                x[q].x = a1 * b2 - a2 * b1;                  // c = a x b (cross1)
                x[q].y = a2 * b0 - a0 * b2;
                park_r[2 * (q * NT + threadIdx.x)] = a0 * b1 - a1 * b0;   // third component, own slot
            }
            fft_line<T, M, E, -1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
            unpack_store_pair<T, M, E, SYNC, NBUF>(x, out + 0 * a.out_fs + line * a.out_ls,
                                                   out + 1 * a.out_fs + line * a.out_ls, t, nk, a.scale, sm, map,
                                                   BUFSTRIDE, phase);
#pragma unroll
            for (int q = 0; q < E; ++q) { x[q].x = park_r[2 * (q * NT + threadIdx.x)]; x[q].y = (T)0; }
            fft_line<T, M, E, -1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
            if (valid) {
                V* C = out + 2 * a.out_fs + line * a.out_ls;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const int k = t + q * P;
                    if (k < nk) C[k] = cscale<T>(x[q], a.scale);
                }
            }
        } else if (MODE == Z_DOT) {
            // sum_j a_j b_j with a = (f0,f1,f2), b = (f3,f4,f5): u_j du_i/dx_j (NS.py:138-145) -> one field
            V p01[E], p23[E], p45[E];
            load_pair<T, M, E>(p01, in + 0 * a.in_fs + line * a.in_ls, in + 1 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p01, t, a.tw, sm, map, BUFSTRIDE, phase);
            load_pair<T, M, E>(p23, in + 2 * a.in_fs + line * a.in_ls, in + 3 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p23, t, a.tw, sm, map, BUFSTRIDE, phase);
            load_pair<T, M, E>(p45, in + 4 * a.in_fs + line * a.in_ls, in + 5 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p45, t, a.tw, sm, map, BUFSTRIDE, phase);
#pragma unroll
            for (int q = 0; q < E; ++q) {
                p01[q].x = p01[q].x * p23[q].y + p01[q].y * p45[q].x + p23[q].x * p45[q].y;
                p01[q].y = (T)0;
            }
            fft_line<T, M, E, -1, SYNC, NBUF>(p01, t, a.tw, sm, map, BUFSTRIDE, phase);
            if (valid) {
                V* C = out + line * a.out_ls;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const int k = t + q * P;
                    if (k < nk) C[k] = cscale<T>(p01[q], a.scale);
                }
            }
        } else if (MODE == Z_UU) {
            // three fields u -> six products u_i u_j, i <= j (divergence_convection, NS.py:147-162)
            V p01[E], p2[E];
            load_pair<T, M, E>(p01, in + 0 * a.in_fs + line * a.in_ls, in + 1 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p01, t, a.tw, sm, map, BUFSTRIDE, phase);
            load_pair<T, M, E>(p2, in + 2 * a.in_fs + line * a.in_ls, (const V*)nullptr, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p2, t, a.tw, sm, map, BUFSTRIDE, phase);
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
                V x[E];
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const T u0 = p01[q].x, u1 = p01[q].y, u2 = p2[q].x;
                    // output order: 00 01 | 02 11 | 12 22
                    x[q].x = pr == 0 ? u0 * u0 : (pr == 1 ? u0 * u2 : u1 * u2);
                    x[q].y = pr == 0 ? u0 * u1 : (pr == 1 ? u1 * u1 : u2 * u2);
                }
                fft_line<T, M, E, -1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
                unpack_store_pair<T, M, E, SYNC, NBUF>(x, out + (2 * pr) * a.out_fs + line * a.out_ls,
                                                       out + (2 * pr + 1) * a.out_fs + line * a.out_ls,
                                                       t, nk, a.scale, sm, map, BUFSTRIDE, phase);
            }
        } else if (MODE == Z_NS2D || MODE == Z_BQ2D) {
            // 2-D solvers: the contiguous axis of a doubly periodic grid.  NS2D (solvers/NS2D.py:40-48): fields
            // (u0, u1, curl) -> (u1*curl, -u0*curl).  Bq2D (solvers/Bq2D.py:121-137): fields (u0, u1, rho, curl) ->
            // (u1*curl, -u0*curl, u0*rho, u1*rho).
            V p01[E], p23[E];
            load_pair<T, M, E>(p01, in + 0 * a.in_fs + line * a.in_ls, in + 1 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p01, t, a.tw, sm, map, BUFSTRIDE, phase);
            if (MODE == Z_BQ2D)        // second pair = curl + i rho
                load_pair<T, M, E>(p23, in + 3 * a.in_fs + line * a.in_ls, in + 2 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            else                       // curl alone
                load_pair<T, M, E>(p23, in + 2 * a.in_fs + line * a.in_ls, (const V*)nullptr, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p23, t, a.tw, sm, map, BUFSTRIDE, phase);
            {
                V x[E];
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const T u0 = p01[q].x, u1 = p01[q].y, w = p23[q].x;
                    x[q].x = u1 * w; x[q].y = -u0 * w;
                }
                fft_line<T, M, E, -1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
                unpack_store_pair<T, M, E, SYNC, NBUF>(x, out + 0 * a.out_fs + line * a.out_ls, out + 1 * a.out_fs + line * a.out_ls,
                                                       t, nk, a.scale, sm, map, BUFSTRIDE, phase);
            }
            if (MODE == Z_BQ2D) {
                V x[E];
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const T u0 = p01[q].x, u1 = p01[q].y, r = p23[q].y;
                    x[q].x = u0 * r; x[q].y = u1 * r;
                }
                fft_line<T, M, E, -1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
                unpack_store_pair<T, M, E, SYNC, NBUF>(x, out + 2 * a.out_fs + line * a.out_ls, out + 3 * a.out_fs + line * a.out_ls,
                                                       t, nk, a.scale, sm, map, BUFSTRIDE, phase);
            }
        } else {
            // MHD (MHD.py:119-127, 99-110): u = (p01.x,p01.y,p23.x), b = (p23.y,p45.x,p45.y)
            V p01[E], p23[E], p45[E];
            load_pair<T, M, E>(p01, in + 0 * a.in_fs + line * a.in_ls, in + 1 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p01, t, a.tw, sm, map, BUFSTRIDE, phase);
            load_pair<T, M, E>(p23, in + 2 * a.in_fs + line * a.in_ls, in + 3 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p23, t, a.tw, sm, map, BUFSTRIDE, phase);
            load_pair<T, M, E>(p45, in + 4 * a.in_fs + line * a.in_ls, in + 5 * a.in_fs + line * a.in_ls, t, a.nin_keep);
            fft_line<T, M, E, +1, SYNC, NBUF>(p45, t, a.tw, sm, map, BUFSTRIDE, phase);
            // z0 = u + b, z1 = u - b ; ZZ[i][j] = F(z0_i * z1_j)
            T z0[3][E], z1[3][E];
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const T u0 = p01[q].x, u1 = p01[q].y, u2 = p23[q].x;
                const T b0 = p23[q].y, b1 = p45[q].x, b2 = p45[q].y;
                z0[0][q] = u0 + b0; z0[1][q] = u1 + b1; z0[2][q] = u2 + b2;
                z1[0][q] = u0 - b0; z1[1][q] = u1 - b1; z1[2][q] = u2 - b2;
            }
#pragma unroll
            for (int pr = 0; pr < 5; ++pr) {
                const int fa = 2 * pr, fb = 2 * pr + 1;
                V x[E];
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    x[q].x = z0[fa / 3][q] * z1[fa % 3][q];
                    x[q].y = fb < 9 ? z0[fb / 3][q] * z1[fb % 3][q] : (T)0;
                }
                fft_line<T, M, E, -1, SYNC, NBUF>(x, t, a.tw, sm, map, BUFSTRIDE, phase);
                unpack_store_pair<T, M, E, SYNC, NBUF>(x, out + fa * a.out_fs + line * a.out_ls,
                                                       fb < 9 ? out + fb * a.out_fs + line * a.out_ls : nullptr,
                                                       t, nk, a.scale, sm, map, BUFSTRIDE, phase);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------
// Fused z-pass for lines that live in ONE warp (P = M/E = 32), NS / VV cross product.
//   * every spectral element is read from HBM exactly once: thread t keeps the raw modes
//     kk = t + 32 q' (q' < QN); the Hermitian mirror A[M-k] needed by the two-lines-per-FFT
//     packing sits in lane (32-t)%32 and comes over with a warp shuffle, as does the mirror of
//     the r2c unpack -- no second global read, no shared-memory round trip;
//   * persistent warps, software pipelined: the loads of the next pair (or of the next line's
//     first pair) are in flight while the current pair is transformed;
//   * two of the three real-space pairs are parked in thread-private shared-memory slots.
// ---------------------------------------------------------------------------------------
template <typename V> __device__ __forceinline__ V shfl_c(V v, int src) {
    V r; r.x = __shfl_sync(0xffffffffu, v.x, src); r.y = __shfl_sync(0xffffffffu, v.y, src); return r;
}

template <typename T, int QN, typename V>
__device__ __forceinline__ void zx_load_raw(V (&raw)[2 * QN], const V* __restrict__ A, const V* __restrict__ B,
                                            int t, int nkeep) {
#pragma unroll
    for (int q = 0; q < QN; ++q) {
        const int kk = t + 32 * q;
        if (kk < nkeep) { raw[2 * q] = A[kk]; raw[2 * q + 1] = B[kk]; }
        else { raw[2 * q] = czero<V>(); raw[2 * q + 1] = czero<V>(); }
    }
}

// x[q] = Za[k] + i Zb[k], k = t + 32 q, from the raw half spectra (see load_pair for the algebra)
template <typename T, int E, int QN, typename V>
__device__ __forceinline__ void zx_build(V (&x)[E], const V (&raw)[2 * QN], int t) {
    const int src = (32 - t) & 31;
#pragma unroll
    for (int q = 0; q < E; ++q) {
        V va = czero<V>(), vb = czero<V>();
        if (q < E / 2) {
            if (q < QN) { va = raw[2 * q]; vb = raw[2 * q + 1]; }
        } else {
            // mirror: k > M/2 -> element M-k = (32-t) + 32 (E-1-q), held by lane (32-t)%32 in slot E-1-q;
            // lane 0 holds its own mirror 32 (E-q) in slot E-q
            constexpr int dummy = 0; (void)dummy;
            const int sq = E - 1 - q;
            V ma = czero<V>(), mb = czero<V>();
            if (sq < QN) { ma = shfl_c(raw[2 * (sq < QN ? sq : 0)], src); mb = shfl_c(raw[2 * (sq < QN ? sq : 0) + 1], src); }
            V oa = czero<V>(), ob = czero<V>();
            if (E - q < QN) { oa = raw[2 * (E - q < QN ? E - q : 0)]; ob = raw[2 * (E - q < QN ? E - q : 0) + 1]; }
            if (q == E / 2) {
                if (t == 0) { va = oa; vb = ob; }               // k = M/2: direct, slot E/2
                else { va = cconj(ma); vb = cconj(mb); }
            } else {
                va = cconj(t == 0 ? oa : ma); vb = cconj(t == 0 ? ob : mb);
            }
        }
        if (t == 0 && (q == 0 || q == E / 2)) { va.y = 0; vb.y = 0; }   // c2r ignores Im of DC / Nyquist
        x[q].x = va.x - vb.y; x[q].y = va.y + vb.x;
    }
}

// r2c unpack with the mirror from shuffles: x = FFT(c + i d); stores C[k], D[k] for k <= M/2, k < nk
template <typename T, int E, typename V>
__device__ __forceinline__ void zx_unpack_store(const V (&x)[E], V* __restrict__ C, V* __restrict__ D,
                                                int t, int nk, T s) {
    const int src = (32 - t) & 31;
    const T h = (T)0.5 * s;
#pragma unroll
    for (int q = 0; q <= E / 2; ++q) {
        V zm;
        if (q < E / 2) {
            const V sh = shfl_c(x[E - 1 - q], src);
            zm = (t == 0) ? x[(E - q) % E] : sh;
        } else {
            zm = x[E / 2];                                   // only lane 0 stores k = M/2
        }
        const int k = t + 32 * q;
        if ((q < E / 2 || t == 0) && k < nk) {
            V c, d;
            c.x = h * (x[q].x + zm.x); c.y = h * (x[q].y - zm.y);
            d.x = h * (x[q].y + zm.y); d.y = -h * (x[q].x - zm.x);
            C[k] = c;
            D[k] = d;
        }
    }
}

template <typename T, int M, int E, int LPC, int QN, int MINB>
__global__ void __launch_bounds__(32 * LPC, MINB)
zx_kernel(const ZArgs<T> a) {
    typedef typename C2<T>::type V;
    static_assert(M / E == 32, "one warp per line");
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int PADW = 128 / (int)sizeof(V);
    constexpr int LP = M + M / PADW + 1;
    const int t = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    V* ex = sm + w * LP;                                   // this warp's exchange line
    V* park = sm + LPC * LP + w * (2 * E * 32);            // this warp's parking slots [2E][32]
    T* park_r = reinterpret_cast<T*>(park);
    SmemLine<1, PADW> map; map.base = 0;
    int phase = 0;
    const V* in = reinterpret_cast<const V*>(a.in);
    V* out = reinterpret_cast<V*>(a.out);
    const long long stride = (long long)gx * LPC;
    long long line = (long long)bx * LPC + w;
    V nxt[2 * QN];
    if (line < a.nlines)
        zx_load_raw<T, QN>(nxt, in + line * a.in_ls, in + a.in_fs + line * a.in_ls, t, a.nin_keep);
    for (; line < a.nlines; line += stride) {
        V x[E];
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
            V raw[2 * QN];
#pragma unroll
            for (int i = 0; i < 2 * QN; ++i) raw[i] = nxt[i];
            if (pr < 2) {
                zx_load_raw<T, QN>(nxt, in + (2 * pr + 2) * a.in_fs + line * a.in_ls,
                                   in + (2 * pr + 3) * a.in_fs + line * a.in_ls, t, a.nin_keep);
            } else if (line + stride < a.nlines) {
                zx_load_raw<T, QN>(nxt, in + (line + stride) * a.in_ls, in + a.in_fs + (line + stride) * a.in_ls,
                                   t, a.nin_keep);
            }
            zx_build<T, E, QN>(x, raw, t);
            fft_line<T, M, E, +1, 1, 1>(x, t, a.tw, ex, map, 0, phase);
            if (pr < 2) {
#pragma unroll
                for (int q = 0; q < E; ++q) park[(pr * E + q) * 32 + t] = x[q];
            }
        }
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const V p01 = park[q * 32 + t], p23 = park[(E + q) * 32 + t];
            const T a0 = p01.x, a1 = p01.y, a2 = p23.x;
            const T b0 = p23.y, b1 = x[q].x, b2 = x[q].y;
            x[q].x = a1 * b2 - a2 * b1;                      // c = a x b (cross1)
            x[q].y = a2 * b0 - a0 * b2;
            park_r[2 * (q * 32 + t)] = a0 * b1 - a1 * b0;
        }
        fft_line<T, M, E, -1, 1, 1>(x, t, a.tw, ex, map, 0, phase);
        zx_unpack_store<T, E>(x, out + line * a.out_ls, out + a.out_fs + line * a.out_ls, t, a.nout_keep, a.scale);
#pragma unroll
        for (int q = 0; q < E; ++q) { x[q].x = park_r[2 * (q * 32 + t)]; x[q].y = (T)0; }
        fft_line<T, M, E, -1, 1, 1>(x, t, a.tw, ex, map, 0, phase);
        V* C = out + 2 * a.out_fs + line * a.out_ls;
#pragma unroll
        for (int q = 0; q <= E / 2; ++q) {
            const int k = t + 32 * q;
            if ((q < E / 2 || t == 0) && k < a.nout_keep) C[k] = cscale<T>(x[q], a.scale);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// zx_kernel with its input staged by the TMA unit: lane 0 of every warp copies the six spectral lines of the NEXT
// line of the warp into a shared-memory stage with bulk-async copies (cp.async.bulk, mbarrier-tracked; the lines are
// contiguous K2p-element segments) while the warp transforms the current one.  The raw modes and their Hermitian
// mirrors are then read from the stage: no prefetch registers (zx_kernel holds 4 QN complex values for them), no
// shuffles on the c2r side, and the loads are a full line ahead instead of one pair-transform.
// NST = 2: two stages per warp (the next line is requested before the current one is touched);
// NST = 1: one stage, refilled as soon as the third pair of the current line has been built from it.
// ---------------------------------------------------------------------------------------
template <typename T, int M, int E, int LPC, int QN, int NST, int MINB>
__global__ void __launch_bounds__(32 * LPC, MINB)
zb_kernel(const ZArgs<T> a) {
    typedef typename C2<T>::type V;
    static_assert(M / E == 32, "one warp per line");
    static_assert(LPC * NST <= 16, "barrier block");
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    constexpr int PADW = 128 / (int)sizeof(V);
    constexpr int LP = (M + M / PADW + 2) & ~1;            // exchange line, even length (keeps the stage 16-byte aligned)
    constexpr int SL = 32 * QN;                            // elements per staged field (>= the line pitch K2p)
    constexpr int WSZ = LP + 2 * E * 32 + NST * 6 * SL;    // elements per warp
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smraw);
    V* sm = reinterpret_cast<V*>(smraw + 128);
    const int t = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    V* ex = sm + w * WSZ;                                  // this warp's exchange line
    V* park = ex + LP;                                     // parking slots [2E][32]
    V* stage = park + 2 * E * 32;                          // [NST][6][SL]
    T* park_r = reinterpret_cast<T*>(park);
    unsigned long long* bar = bars + w * NST;
    SmemLine<1, PADW> map; map.base = 0;
    int phase = 0;
    const V* in = reinterpret_cast<const V*>(a.in);
    V* out = reinterpret_cast<V*>(a.out);
    const long long stride = (long long)gx * LPC;
    long long line = (long long)bx * LPC + w;
    const unsigned int bytes = (unsigned int)(a.in_ls * (long long)sizeof(V));
    auto issue = [&](long long ln, int s) {                // lane 0 only
        mbar_expect_tx(&bar[s], 6u * bytes);
#pragma unroll
        for (int f = 0; f < 6; ++f) bulk_g2s(stage + (s * 6 + f) * SL, in + f * a.in_fs + ln * a.in_ls, bytes, &bar[s]);
    };
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
        mbar_init_fence();
        if (line < a.nlines) issue(line, 0);
    }
    __syncwarp();
    for (int n = 0; line < a.nlines; line += stride, ++n) {
        const int s = NST == 2 ? (n & 1) : 0;
        if (NST == 2 && t == 0 && line + stride < a.nlines) issue(line + stride, s ^ 1);
        mbar_wait(&bar[s], (unsigned int)((NST == 2 ? (n >> 1) : n) & 1));
        V x[E];
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
            const V* A = stage + (s * 6 + 2 * pr) * SL;
            const V* B = A + SL;
            // x[q] = Za[k] + i Zb[k], k = t + 32 q (load_pair's algebra on the staged half spectra)
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const int k = t + 32 * q;
                const bool lowh = q < E / 2 || (q == E / 2 && t == 0);         // 2 k <= M
                const int kk = lowh ? k : M - k;
                V va = czero<V>(), vb = czero<V>();
                if (kk < a.nin_keep) { va = A[kk]; vb = B[kk]; }
                if (!lowh) { va.y = -va.y; vb.y = -vb.y; }
                if (t == 0 && (q == 0 || q == E / 2)) { va.y = 0; vb.y = 0; }   // c2r ignores Im of DC / Nyquist
                x[q].x = va.x - vb.y; x[q].y = va.y + vb.x;
            }
            if (NST == 1 && pr == 2) {
                __syncwarp();                                   // every lane has read the stage
                if (t == 0 && line + stride < a.nlines) issue(line + stride, 0);
            }
            fft_line<T, M, E, +1, 1, 1>(x, t, a.tw, ex, map, 0, phase);
            if (pr < 2) {
#pragma unroll
                for (int q = 0; q < E; ++q) park[(pr * E + q) * 32 + t] = x[q];
            }
        }
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const V p01 = park[q * 32 + t], p23 = park[(E + q) * 32 + t];
            const T a0 = p01.x, a1 = p01.y, a2 = p23.x;
            const T b0 = p23.y, b1 = x[q].x, b2 = x[q].y;
            x[q].x = a1 * b2 - a2 * b1;                      // c = a x b (cross1)
            x[q].y = a2 * b0 - a0 * b2;
            park_r[2 * (q * 32 + t)] = a0 * b1 - a1 * b0;
        }
        fft_line<T, M, E, -1, 1, 1>(x, t, a.tw, ex, map, 0, phase);
        zx_unpack_store<T, E>(x, out + line * a.out_ls, out + a.out_fs + line * a.out_ls, t, a.nout_keep, a.scale);
#pragma unroll
        for (int q = 0; q < E; ++q) { x[q].x = park_r[2 * (q * 32 + t)]; x[q].y = (T)0; }
        fft_line<T, M, E, -1, 1, 1>(x, t, a.tw, ex, map, 0, phase);
        V* C = out + 2 * a.out_fs + line * a.out_ls;
#pragma unroll
        for (int q = 0; q <= E / 2; ++q) {
            const int k = t + 32 * q;
            if ((q < E / 2 || t == 0) && k < a.nout_keep) C[k] = cscale<T>(x[q], a.scale);
        }
        __syncwarp();                                           // the stage this iteration read may be refilled
    }
}

// ---------------------------------------------------------------------------------------
// Fused z-pass, one CTA of P = M/E threads (two or four warps) per line: the zx_kernel scheme for the long lines
// whose register budget does not allow one warp per line (fp64 M >= 512, fp32 M >= 2048).  Same single HBM read
// per spectral element, same software pipeline; the Hermitian mirrors travel through the (idle) exchange line in
// shared memory instead of warp shuffles.
// ---------------------------------------------------------------------------------------
template <typename T, int M, int E, int QN, int MINB>
__global__ void __launch_bounds__(M / E, MINB)
zy_kernel(const ZArgs<T> a) {
    typedef typename C2<T>::type V;
    constexpr int P = M / E;
    static_assert(P == 64 || P == 128, "two or four warps per line");
    SDNS_DYN_SMEM(smraw);
    SDNS_XFER_ROLE(a, smraw)
    V* sm = reinterpret_cast<V*>(smraw);
    constexpr int PADW = 128 / (int)sizeof(V);
    constexpr int LP = M + M / PADW + 1;
    constexpr int EXN = LP > 2 * QN * P ? LP : 2 * QN * P;
    const int t = threadIdx.x;
    const int src = (P - t) % P;
    V* ex = sm;                                            // exchange line; doubles as mirror staging [slot][thread]
    V* park = sm + EXN;                                    // parking slots [2E][P]
    T* park_r = reinterpret_cast<T*>(park);
    SmemLine<1, PADW> map; map.base = 0;
    int phase = 0;
    const V* in = reinterpret_cast<const V*>(a.in);
    V* out = reinterpret_cast<V*>(a.out);
    const long long stride = gx;
    long long line = bx;
    V nxt[2 * QN];
    auto load_raw = [&](V (&r)[2 * QN], const V* __restrict__ A, const V* __restrict__ B) {
#pragma unroll
        for (int q = 0; q < QN; ++q) {
            const int kk = t + P * q;
            if (kk < a.nin_keep) { r[2 * q] = A[kk]; r[2 * q + 1] = B[kk]; }
            else { r[2 * q] = czero<V>(); r[2 * q + 1] = czero<V>(); }
        }
    };
    if (line < a.nlines) load_raw(nxt, in + line * a.in_ls, in + a.in_fs + line * a.in_ls);
    for (; line < a.nlines; line += stride) {
        V x[E];
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
            V raw[2 * QN];
#pragma unroll
            for (int i = 0; i < 2 * QN; ++i) raw[i] = nxt[i];
            if (pr < 2) load_raw(nxt, in + (2 * pr + 2) * a.in_fs + line * a.in_ls, in + (2 * pr + 3) * a.in_fs + line * a.in_ls);
            else if (line + stride < a.nlines) load_raw(nxt, in + (line + stride) * a.in_ls, in + a.in_fs + (line + stride) * a.in_ls);
            __syncthreads();                               // the previous transform has left the exchange line
#pragma unroll
            for (int i = 0; i < 2 * QN; ++i) ex[i * P + t] = raw[i];
            __syncthreads();
            // x[q] = Za[k] + i Zb[k], k = t + P q (load_pair's algebra, mirrors from the staging)
#pragma unroll
            for (int q = 0; q < E; ++q) {
                V va = czero<V>(), vb = czero<V>();
                if (q < E / 2) {
                    if (q < QN) { va = raw[2 * q]; vb = raw[2 * q + 1]; }
                } else {
                    const int sq = E - 1 - q;              // k > M/2: element M-k sits in thread (P-t)%P, slot E-1-q ...
                    V ma = czero<V>(), mb = czero<V>();
                    if (sq < QN) { ma = ex[(2 * sq) * P + src]; mb = ex[(2 * sq + 1) * P + src]; }
                    V oa = czero<V>(), ob = czero<V>();    // ... thread 0 holds its own mirror P (E-q) in slot E-q
                    if (E - q < QN) { oa = raw[2 * (E - q < QN ? E - q : 0)]; ob = raw[2 * (E - q < QN ? E - q : 0) + 1]; }
                    if (q == E / 2) {
                        if (t == 0) { va = oa; vb = ob; }  // k = M/2: direct
                        else { va = cconj(ma); vb = cconj(mb); }
                    } else {
                        va = cconj(t == 0 ? oa : ma); vb = cconj(t == 0 ? ob : mb);
                    }
                }
                if (t == 0 && (q == 0 || q == E / 2)) { va.y = 0; vb.y = 0; }
                x[q].x = va.x - vb.y; x[q].y = va.y + vb.x;
            }
            fft_line<T, M, E, +1, 0, 1>(x, t, a.tw, ex, map, 0, phase);
            if (pr < 2) {
#pragma unroll
                for (int q = 0; q < E; ++q) park[(pr * E + q) * P + t] = x[q];
            }
        }
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const V p01 = park[q * P + t], p23 = park[(E + q) * P + t];
            const T a0 = p01.x, a1 = p01.y, a2 = p23.x;
            const T b0 = p23.y, b1 = x[q].x, b2 = x[q].y;
            x[q].x = a1 * b2 - a2 * b1;                      // c = a x b (cross1)
            x[q].y = a2 * b0 - a0 * b2;
            park_r[2 * (q * P + t)] = a0 * b1 - a1 * b0;
        }
        fft_line<T, M, E, -1, 0, 1>(x, t, a.tw, ex, map, 0, phase);
        // r2c unpack of the pair (c0, c1): mirrors through the staging
        __syncthreads();
#pragma unroll
        for (int q = 0; q < E; ++q) ex[q * P + t] = x[q];
        __syncthreads();
        {
            const T h = (T)0.5 * a.scale;
            V* C = out + line * a.out_ls;
            V* D = out + a.out_fs + line * a.out_ls;
#pragma unroll
            for (int q = 0; q <= E / 2; ++q) {
                V zm;
                if (q < E / 2) zm = (t == 0) ? x[(E - q) % E] : ex[(E - 1 - q) * P + src];
                else zm = x[E / 2];
                const int k = t + P * q;
                if ((q < E / 2 || t == 0) && k < a.nout_keep) {
                    V c, d;
                    c.x = h * (x[q].x + zm.x); c.y = h * (x[q].y - zm.y);
                    d.x = h * (x[q].y + zm.y); d.y = -h * (x[q].x - zm.x);
                    C[k] = c; D[k] = d;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < E; ++q) { x[q].x = park_r[2 * (q * P + t)]; x[q].y = (T)0; }
        fft_line<T, M, E, -1, 0, 1>(x, t, a.tw, ex, map, 0, phase);
        V* C2p = out + 2 * a.out_fs + line * a.out_ls;
#pragma unroll
        for (int q = 0; q <= E / 2; ++q) {
            const int k = t + P * q;
            if ((q < E / 2 || t == 0) && k < a.nout_keep) C2p[k] = cscale<T>(x[q], a.scale);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Hermitian-weighted sum |u_hat|^2 (shenfun.fourier.energy_fourier as used by tests/TG.py:101)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void energy_kernel(const typename C2<T>::type* __restrict__ u, long long n, int Nh, int N2,
                              double* __restrict__ out) {
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int k2 = (int)(i % Nh);
        const typename C2<T>::type v = u[i];
        const double w = (k2 == 0 || 2 * k2 == N2) ? 1.0 : 2.0;
        s += w * ((double)v.x * v.x + (double)v.y * v.y);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    SDNS_STATIC_SMEM(double, ws, 32);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[blockIdx.x] = s;
    }
}

}  // namespace sdns
