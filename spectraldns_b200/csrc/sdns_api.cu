// sdns_api.cu -- plan object and C ABI of libsdns_b200.so (see include/sdns_b200.h).
//
// Host-side equivalent of the reference's get_context() setup (solvers/NS.py:12-72) and of the
// call sequence ComputeRHS -> conv -> transforms (NS.py:191-261, VV.py:92-146, MHD.py:119-176),
// re-expressed as five kernel launches per right-hand side (B0, B1, Z, F1, F0).
#ifndef SDNS_HOST_SHIM
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include <deque>
#include "../../include/sdns_b200.h"
#include "launch.cuh"

using namespace sdns;

// grids of the grid-stride elementwise / reduction kernels: 8 CTAs per SM of a B200 (the emulated build keeps them tiny)
#ifdef SDNS_HOST_SHIM
#define SDNS_EW_BLOCKS 2
#define SDNS_RED_BLOCKS 2
#else
#define SDNS_EW_BLOCKS 1184
#define SDNS_RED_BLOCKS 1024
#endif

#define SDNS_MAX_BINS 4096        // shells of sdns_spectrum (a 4096^3 grid has 2048)

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    return fail(SDNS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

typedef int (*launch_fn)(int, const void*, cudaStream_t);
extern launch_fn g_launch[FAM_COUNT][2];
launch_fn g_launch[FAM_COUNT][2] = {       // also used by sdns2d_api.cu
#define ROW(f) { sdns_launch_##f##_f32, sdns_launch_##f##_f64 },
    ROW(0) ROW(1) ROW(2) ROW(3) ROW(4) ROW(5) ROW(6) ROW(7) ROW(8) ROW(9) ROW(10)
    ROW(11) ROW(12) ROW(13) ROW(14) ROW(15) ROW(16)
#undef ROW
};

static bool size_ok(int n) {
    switch (n) {
#define X(N) case N: return true;
        SDNS_SIZES(X)
        SDNS_SIZES_5(X)
#undef X
        default: return false;
    }
}

// geometry of one function space (T or Tp)
struct Space {
    int M[3];              // transform lengths (physical shape)
    AxisMap bmap[3];       // backward-input map per axis (transform index -> memory index)
    AxisMap fmap[2];       // forward-output map, axes 0,1
    int K1n;               // axis-1 modes entering the backward transform (compact count)
    int col_nlo, col_gap;  // compact axis-1 index -> memory index
    int K2n, K2p;          // axis-2 modes entering the backward transform, padded pitch
    double scale;          // 1/prod(M)
    // slab decomposition: this rank owns the spectral k1 in [s1, s1 + N1l) (SDNS_K1_BLOCKS) or the modes rank, rank + P,
    // ... (SDNS_K1_CYCLIC) and the physical x0 in [s0, s0 + M0l)  (spectralDNS3D_short.py:28-29); s = rank * share when
    // the rank count divides the extent, else mpi4py-fft's uneven shares (share_count / share_start)
    int M0l;               // local physical planes (this rank's share of M[0]: M0/P + (rank < M0 % P), spectralinit's split)
    int M0p, s0;           // ceil(M0/P): the plane pitch of every rank's W0 and send slots; first plane of this rank
    size_t xtab_off;       // uneven split of M[0]: B0 store table, x0 -> (owner << 24 | local plane)
    int K1l;               // kept axis-1 modes owned by this rank
    int lcol_nlo, lcol_gap;// local compact axis-1 index -> local memory index
    int c1off;             // position of this rank's first kept axis-1 mode in W0's compact axis 1 (rank-major order)
    size_t itab_off, otab_off;   // CYCLIC: B1 row table (transform index -> W0 row) and F1 store table (-> rank, row)
    int cyc_first[8], cyc_hi[8], cyc_dm;   // the same map in closed form (StridedArgs, passes.cuh)
};

struct sdns_plan {
    sdns_config cfg;
    int N[3], Nh, Nhp;
    int P, rank, N1l;       // ranks, this rank, local spectral extent of axis 1 (N1/P + (rank < N1 % P))
    int N1p, s1;            // ceil(N1/P): the row pitch of every rank's W3 and send slots; first axis-1 mode of this rank
    int uneven1;            // N1 % P != 0: the axis-1 forward pass finds (owner, row) in a table
    int k1cyc;              // 1: SDNS_K1_CYCLIC (local index j holds mode rank + j*P), 0: SDNS_K1_BLOCKS
    size_t off_C, bytes_C, off_flags, off_D, bytes_D, off_S, bytes_S, off_U, bytes_U;
    bool own_ws;            // workspace cudaMalloc'ed by the library (multi-GPU: IPC-shared)
    char* peer_ws[8];       // base of every rank's workspace (peer_ws[rank] == ws)
    unsigned int epoch;
    int prec;               // 0 f32, 1 f64
    size_t rs, cs;          // sizeof real / complex
    Space sp[2];            // SDNS_SPACE_T, SDNS_SPACE_TP
    cudaStream_t stream;
    // workspace carving (byte offsets)
    size_t ws_need, off_tab, off_A, off_B, off_red, bytes_A, bytes_B;
    char* ws; size_t ws_bytes;
    std::vector<char> host_tables;                 // image of the table region
    std::map<int, size_t> tw_off;                  // transform length -> offset of its twiddles
    size_t kx_off, ky_off, kz_off;
    long long launches;
    int red_blocks;
    // multi-GPU exchange.  xmode 0: B0 / F1 store straight into the peers (one launch per pass, serial schedule).
    // xmode 1: B0 / F1 write per-destination send buffers in chunks and the copy engines move each chunk over
    // NVLink (one stream per peer) while the SMs work on the next chunk.
    int xmode, nchunk, nsplit;      // nsplit: streams (copy engines) per destination rank
    // xmode 2 (default): the send slots are moved by a transfer ROLE inside the following pass kernels (xfer.cuh):
    // pending 2-D copies wait in `pend`; every launch of the plan stream takes rows worth xratio x its own HBM bytes.
    std::deque<XferBatch> pend;
    int xctas, xflush_ctas, xtma;   // transfer CTAs per carrying launch / of a transfer-only launch; bulk-async or ld/st
    double xratio;                  // NVLink bytes a launch carries per byte of its own HBM traffic
    unsigned int xinflight;         // ring bytes per carrying launch (sets its number of transfer CTAs); 0: xctas as given
    double xfer_bytes; long long xfer_flushes;
    int kcopy, kcopy_ctas;          // xmode 1 variant: a grid-capped copy kernel (peer stores) instead of cudaMemcpy2DAsync
    // SDNS_GRAPH=1 (experiment, multi-GPU copy-engine mode): an RK4 step -- kernels, peer copies and their cross-stream
    // edges -- is captured once per argument set into a CUDA graph on the library's own stream and replayed, which
    // takes the host out of the schedule.  Barriers inside a graph take their epoch from a device-resident counter.
    int use_graph; bool capturing, gwarm, ghave;
    cudaStream_t gstream; cudaEvent_t ev_gin, ev_gout; cudaGraphExec_t gexec;
    struct GKey { void* u; void* u1; void* u2; const void* src; double dt, nu, eta; } gkey;
    long long glaunches;
    std::vector<cudaStream_t> ys;   // copy streams (a few, shared by the destinations) x nsplit parts
    std::vector<cudaEvent_t> ev_k;  // [nchunk] the pass of chunk c has finished
    std::vector<cudaEvent_t> ev_y;  // per copy stream: drained
    size_t off_SF, bytes_SF;        // F1 send buffers (B0's live in the W1 buffer, which is idle at that point)
    bool b0_preissued;              // the B0 chunks (and copies) of the coming right-hand side are already enqueued ...
    const void* b0_pre_u;           // ... for this input array (in the work layout); honoured only for exactly that call
    struct CRec { int s; cudaEvent_t a, b; double bytes; };
    std::vector<CRec> crecs;
    double copy_ms[32]; double copy_bytes; long long copy_n;
    // optional per-family profiling (bench.py roofline): CUDA events around every launch
    bool prof;
    std::vector<cudaEvent_t> ev_pool; size_t ev_used;
    struct Rec { int fam; cudaEvent_t a, b; double bytes, remote; };
    std::vector<Rec> recs;
    // timeline (sdns_profile_enable(plan, 2)): start / end of every launch, copy and barrier relative to tl_base
    bool tl_on; cudaEvent_t tl_base;
    std::vector<Rec> brecs;                 // cross-GPU barriers (timeline only)
    std::vector<double> timeline;           // rows of (kind, t0_ms, t1_ms, bytes)
    double prof_ms[FAM_COUNT]; double prof_bytes[FAM_COUNT]; double prof_remote[FAM_COUNT]; long long prof_n[FAM_COUNT];
};

static AxisMap all_map(int n) { AxisMap m; m.nlo = n; m.nhi = 0; m.shift = 0; return m; }

static int default_kcut(int n) { return (int)ceil(2.0 / 3.0 * (n / 2 + 1)) - 1; }

// Share of rank r of an axis of n entries split over P ranks: n/P + (r < n % P), the remainder going to the first ranks
// (mpi4py-fft's decomposition).  start = first index of the share.
static int share_count(int n, int P, int r) { return n / P + (r < n % P ? 1 : 0); }
static int share_start(int n, int P, int r) { return r * (n / P) + std::min(r, n % P); }

// Kept axis-1 modes of rank r in space q: local indices [0, a1) (low run) and [a1 + gap, a1 + gap + nb) (high run).
// The kept global modes are [0, col_nlo) and [N1 - (K1n - col_nlo), N1).
struct K1Own { int a1, nb, gap; };
static K1Own k1_own(const sdns_plan* p, const Space& q, int r) {
    const int N1 = p->N[1], N1l = p->N1l, P = p->P;
    const int nlo = q.col_nlo, nhi = q.K1n - q.col_nlo, hstart = N1 - nhi;
    K1Own o;
    if (!p->k1cyc) {
        const int lo = share_start(N1, P, r), hi = lo + share_count(N1, P, r);
        o.a1 = std::max(0, std::min(hi, nlo) - lo);
        const int b0 = std::max(lo, hstart);
        o.nb = (nhi > 0 && hi > b0) ? hi - b0 : 0;
        o.gap = o.nb > 0 ? (b0 - lo) - o.a1 : 0;
    } else {
        // local index j holds mode j*P + r
        o.a1 = nlo > r ? (nlo - r + P - 1) / P : 0;                     // j*P + r < nlo
        const int jh = hstart > r ? (hstart - r + P - 1) / P : 0;       // first j with j*P + r >= hstart
        o.nb = nhi > 0 ? std::max(0, N1l - jh) : 0;
        o.gap = o.nb > 0 ? jh - o.a1 : 0;
    }
    return o;
}

static void build_spaces(sdns_plan* p) {
    const int* N = p->N;
    // T: plain space
    Space& t = p->sp[SDNS_SPACE_T];
    for (int i = 0; i < 3; ++i) { t.M[i] = N[i]; t.bmap[i] = all_map(i < 2 ? N[i] : p->Nh); }
    t.fmap[0] = all_map(N[0]); t.fmap[1] = all_map(N[1]);
    t.K1n = N[1]; t.col_nlo = N[1]; t.col_gap = 0; t.K2n = p->Nh;
    Space& d = p->sp[SDNS_SPACE_TP];
    d = t;
    if (p->cfg.dealias == SDNS_DEALIAS_23) {
        // truncation of the backward input, |k_i| <= kcut_i (solvers/NS.py:29-31 dealias_direct;
        // cutoff spectralDNS3D_short.py:44-46).  Truncated lines/columns are never transformed.
        for (int i = 0; i < 2; ++i) {
            int kc = p->cfg.kcut[i] >= 0 ? p->cfg.kcut[i] : default_kcut(N[i]);
            if (2 * kc + 1 < N[i]) { d.bmap[i].nlo = kc + 1; d.bmap[i].nhi = kc; d.bmap[i].shift = 0; }
        }
        int kc2 = p->cfg.kcut[2] >= 0 ? p->cfg.kcut[2] : default_kcut(N[2]);
        if (kc2 + 1 < p->Nh) d.K2n = kc2 + 1;
        d.K1n = d.bmap[1].nlo + d.bmap[1].nhi;
        d.col_nlo = d.bmap[1].nlo; d.col_gap = N[1] - d.K1n;
        d.bmap[1].shift = N[1] - d.K1n;            // B1 reads the compact axis-1 layout
    } else if (p->cfg.dealias == SDNS_DEALIAS_32) {
        // zero padding N -> 3N/2 per axis on backward, corner truncation on forward
        for (int i = 0; i < 3; ++i) d.M[i] = (3 * N[i]) / 2;
        for (int i = 0; i < 2; ++i) {
            d.bmap[i].nlo = N[i] / 2; d.bmap[i].nhi = N[i] - N[i] / 2; d.bmap[i].shift = d.M[i] - N[i];
            d.fmap[i] = d.bmap[i];
        }
    }
    for (int s = 0; s < 2; ++s) {
        Space& q = p->sp[s];
        q.K2p = (q.K2n + 1) & ~1;
        q.scale = 1.0 / ((double)q.M[0] * q.M[1] * q.M[2]);
        q.M0l = share_count(q.M[0], p->P, p->rank);
        q.M0p = (q.M[0] + p->P - 1) / p->P;
        q.s0 = share_start(q.M[0], p->P, p->rank);
        q.xtab_off = 0;
        // kept axis-1 modes owned here; W0's compact axis 1 lists the ranks' kept modes one rank after the other
        // (for SDNS_K1_BLOCKS that is the natural order)
        const K1Own o = k1_own(p, q, p->rank);
        q.K1l = o.a1 + o.nb;
        q.lcol_nlo = o.a1;
        q.lcol_gap = o.gap;
        q.c1off = 0;
        for (int r = 0; r < p->rank; ++r) { const K1Own x = k1_own(p, q, r); q.c1off += x.a1 + x.nb; }
        q.itab_off = q.otab_off = 0;
    }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
static void fill_tables(sdns_plan* p) {
    typedef typename C2<T>::type V;
    std::vector<char>& h = p->host_tables;
    auto add_tw = [&](int n) {
        if (p->tw_off.count(n)) return;
        size_t off = align_up(h.size(), 256);
        h.resize(off + sizeof(V) * n);
        V* tw = reinterpret_cast<V*>(h.data() + off);
        for (int j = 0; j < n; ++j) {
            long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
            tw[j].x = (T)cosl(ang); tw[j].y = (T)sinl(ang);
        }
        p->tw_off[n] = off;
    };
    for (int s = 0; s < 2; ++s) for (int i = 0; i < 3; ++i) add_tw(p->sp[s].M[i]);
    auto add_k = [&](int n, int len, double L, bool real_axis, int first = 0, int step = 1) {
        size_t off = align_up(h.size(), 256);
        h.resize(off + sizeof(T) * len);
        T* k = reinterpret_cast<T*>(h.data() + off);
        for (int j = 0; j < len; ++j) {
            const int i = j * step + first;
            int kk = real_axis ? i : (i < (n + 1) / 2 ? i : i - n);
            // same evaluation order as the oracle: k*2*pi/L in double, then cast (NS.py:38-41)
            k[j] = (T)(((double)kk * 2.0 * M_PI) / L);
        }
        return off;
    };
    p->kx_off = add_k(p->N[0], p->N[0], p->cfg.L[0], false);
    // N1p entries on every rank (the table region has the same size everywhere); the entries past this rank's share are unused
    if (p->k1cyc) p->ky_off = add_k(p->N[1], p->N1p, p->cfg.L[1], false, p->rank, p->P);
    else p->ky_off = add_k(p->N[1], p->N1p, p->cfg.L[1], false, p->s1);
    p->kz_off = add_k(p->N[2], p->Nh, p->cfg.L[2], true);
    // SDNS_K1_CYCLIC: the axis-1 passes on either side of a transpose see the modes in rank-major order.  B1 finds the
    // W0 row of transform index j in itab (-1: not an input), F1 finds (owner << 24 | local row) of output j in otab.
    for (int s = 0; s < 2 && p->k1cyc && p->P > 1; ++s) {
        Space& q = p->sp[s];
        const int M1 = q.M[1], P = p->P;
        std::vector<int> first(P, 0);
        std::vector<K1Own> own(P);
        for (int r = 0, c = 0; r < P; ++r) { own[r] = k1_own(p, q, r); first[r] = c; c += own[r].a1 + own[r].nb; }
        for (int r = 0; r < 8; ++r) { q.cyc_first[r] = r < P ? first[r] : 0; q.cyc_hi[r] = r < P ? first[r] - own[r].gap : 0; }
        q.cyc_dm = q.bmap[1].shift - q.col_gap;
        size_t off = align_up(h.size(), 256);
        h.resize(off + sizeof(int) * 2 * M1);
        q.itab_off = off; q.otab_off = off + sizeof(int) * M1;
        std::vector<int> it(M1), ot(M1);
        for (int j = 0; j < M1; ++j) {
            const AxisMap& bm = q.bmap[1];
            it[j] = -1;
            if (j < bm.nlo || j >= M1 - bm.nhi) {
                const int idx = j < bm.nlo ? j : j - bm.shift;                // natural compact index
                const int m = idx < q.col_nlo ? idx : idx + q.col_gap;        // mode index in [0, N1)
                const int r = m % P, jl = m / P;
                it[j] = first[r] + (jl < own[r].a1 ? jl : jl - own[r].gap);
            }
            const AxisMap& fm = q.fmap[1];
            ot[j] = -1;
            if (j < fm.nlo || j >= M1 - fm.nhi) {
                const int i = j < fm.nlo ? j : j - fm.shift;
                ot[j] = ((i % P) << 24) | (i / P);
            }
        }
        memcpy(h.data() + q.itab_off, it.data(), sizeof(int) * M1);
        memcpy(h.data() + q.otab_off, ot.data(), sizeof(int) * M1);
    }
    // Uneven splits (N1 or M0 not divisible by P): the owner of an index is no longer index / chunk, so the pass in front
    // of a transpose finds (owner << 24 | row inside the owner's share) in a table: F1 for its axis-1 outputs, B0 for its x0.
    for (int s = 0; s < 2 && p->P > 1; ++s) {
        Space& q = p->sp[s];
        const int P = p->P;
        auto owner_tab = [&](int n, const std::vector<int>& idx) {      // idx[j]: global index of entry j, or -1
            std::vector<int> t(idx.size());
            for (size_t j = 0; j < idx.size(); ++j) {
                t[j] = -1;
                if (idx[j] < 0) continue;
                int r = 0;
                while (r + 1 < P && share_start(n, P, r + 1) <= idx[j]) ++r;
                t[j] = (r << 24) | (idx[j] - share_start(n, P, r));
            }
            const size_t off = align_up(h.size(), 256);
            h.resize(off + sizeof(int) * t.size());
            memcpy(h.data() + off, t.data(), sizeof(int) * t.size());
            return off;
        };
        if (p->uneven1) {
            const int M1 = q.M[1];
            const AxisMap& fm = q.fmap[1];
            std::vector<int> idx(M1, -1);
            for (int j = 0; j < M1; ++j)
                if (j < fm.nlo || j >= M1 - fm.nhi) idx[j] = j < fm.nlo ? j : j - fm.shift;
            q.otab_off = owner_tab(p->N[1], idx);
        }
        if (q.M[0] % P) {
            std::vector<int> idx(q.M[0]);
            for (int j = 0; j < q.M[0]; ++j) idx[j] = j;
            q.xtab_off = owner_tab(q.M[0], idx);
        }
    }
}

extern "C" int sdns_abi_version(void) { return SDNS_ABI_VERSION; }
extern "C" const char* sdns_last_error(void) { return g_err.c_str(); }
extern "C" int sdns_size_supported(int n, int precision) {
    (void)precision;
    return size_ok(n) ? SDNS_OK : SDNS_ERR_SIZE;
}

extern "C" int sdns_plan_create(sdns_plan** out, const sdns_config* cfg) {
    if (!out || !cfg) return fail(SDNS_ERR_ARG, "null argument");
    if (cfg->abi_version != SDNS_ABI_VERSION) return fail(SDNS_ERR_ARG, "ABI version mismatch");
    if (cfg->precision != SDNS_SINGLE && cfg->precision != SDNS_DOUBLE) return fail(SDNS_ERR_ARG, "precision");
    if (cfg->nranks < 1 || cfg->nranks > 8 || cfg->rank < 0 || cfg->rank >= cfg->nranks)
        return fail(SDNS_ERR_ARG, "rank/nranks: 1..8 ranks (one process per GPU of one node)");
    if (cfg->nranks > 1 && cfg->decomposition != SDNS_SLAB)
        return fail(SDNS_ERR_ARG, "multi-GPU runs use the slab decomposition (pencil is not on the B200 path yet)");
    if (cfg->solver < SDNS_NS || cfg->solver > SDNS_MHD) return fail(SDNS_ERR_ARG, "solver");
    if (cfg->convection < SDNS_CONV_VORTEX || cfg->convection > SDNS_CONV_SKEWED) return fail(SDNS_ERR_ARG, "convection");
    if (cfg->solver == SDNS_VV && cfg->convection != SDNS_CONV_VORTEX)
        return fail(SDNS_ERR_ARG, "VV supports only convection='Vortex' (solvers/VV.py:87-88)");
    if (cfg->solver == SDNS_MHD && cfg->convection != SDNS_CONV_DIVERGENCE)
        return fail(SDNS_ERR_ARG, "MHD supports only convection='Divergence' (solvers/MHD.py:114-115)");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(SDNS_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce) +
                    " (libsdns_b200 has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(cfg->device));
    sdns_plan* p = new sdns_plan();
    p->cfg = *cfg;
    for (int i = 0; i < 3; ++i) {
        p->N[i] = cfg->N[i];
        if (p->N[i] < 2 || p->N[i] % 2) { delete p; return fail(SDNS_ERR_SIZE, "N must be even"); }
        if (!(cfg->L[i] > 0)) { delete p; return fail(SDNS_ERR_ARG, "L must be positive"); }
    }
    p->Nh = p->N[2] / 2 + 1;
    p->Nhp = (p->Nh + 1) & ~1;
    p->P = cfg->nranks; p->rank = cfg->rank;
    p->own_ws = false; p->epoch = 0;
    for (int r = 0; r < 8; ++r) p->peer_ws[r] = nullptr;
    if (p->N[1] < p->P) { delete p; return fail(SDNS_ERR_SIZE, "N[1] must be at least the number of ranks"); }
    // spectral axis 1 split like mpi4py-fft's slabs: N1/P modes per rank, the first N1 % P ranks one more
    p->N1l = share_count(p->N[1], p->P, p->rank);
    p->N1p = (p->N[1] + p->P - 1) / p->P;
    p->s1 = share_start(p->N[1], p->P, p->rank);
    p->uneven1 = p->N[1] % p->P != 0;
    if (cfg->k1_layout != SDNS_K1_BLOCKS && cfg->k1_layout != SDNS_K1_CYCLIC) { delete p; return fail(SDNS_ERR_ARG, "k1_layout"); }
    p->k1cyc = (cfg->k1_layout == SDNS_K1_CYCLIC && p->P > 1) ? 1 : 0;
    if (p->k1cyc && p->uneven1) { delete p; return fail(SDNS_ERR_SIZE, "k1_layout cyclic needs N[1] divisible by the number of ranks"); }
    p->prec = cfg->precision;
    p->rs = p->prec ? 8 : 4; p->cs = 2 * p->rs;
    p->stream = 0; p->ws = nullptr; p->ws_bytes = 0; p->launches = 0;
    p->prof = false; p->ev_used = 0;
    p->tl_on = false; p->tl_base = nullptr;
    p->kcopy = 0; p->kcopy_ctas = 32;
#ifdef SDNS_HOST_SHIM
    p->xctas = 2; p->xflush_ctas = 3;        // every emulated CUDA thread is an OS thread
#else
    p->xctas = 32; p->xflush_ctas = 128;
#endif
    p->xtma = 1; p->xinflight = 2u << 20;
#ifdef SDNS_HOST_SHIM
    p->xinflight = 0;
#endif
    p->xratio = 0.3; p->xfer_bytes = 0; p->xfer_flushes = 0;
    p->use_graph = 0; p->capturing = p->gwarm = p->ghave = false; p->gstream = nullptr; p->ev_gin = p->ev_gout = nullptr; p->gexec = nullptr; p->glaunches = 0;
    p->xmode = 0; p->nchunk = 1; p->nsplit = 1; p->off_SF = 0; p->bytes_SF = 0; p->b0_preissued = false; p->b0_pre_u = nullptr;
    p->copy_bytes = 0; p->copy_n = 0; for (int i = 0; i < 32; ++i) p->copy_ms[i] = 0;
    if (p->P > 1) {
        const char* xm = getenv("SDNS_EXCHANGE");
        // "tma" (default for 3 or more ranks): send slots + transfer role with bulk-async copies; "ldst": the role with plain loads / stores;
        // "ce" / "kcopy": copy engines / copy kernels on side streams (round 1); "store": peer stores fused into the passes
        p->xmode = (xm && !strcmp(xm, "store")) ? 0 : ((xm && (!strcmp(xm, "ce") || !strcmp(xm, "kcopy"))) ? 1 : 2);
        // measured (profiles/r2): with ONE peer the copy engines hide the exchange best (512^3 per GPU: 57.5 ms against
        // 67-70 ms for the transfer role); with seven destinations per chunk the transfer role does (1024^3 on 8 GPUs:
        // 75 ms against 91-96 ms)
        if (!xm && p->P == 2) p->xmode = 1;
        p->xtma = (xm && !strcmp(xm, "ldst")) ? 0 : 1;
        if (const char* v = getenv("SDNS_XCTAS")) { p->xctas = std::max(1, atoi(v)); p->xinflight = 0; }
        if (const char* v = getenv("SDNS_XINFLIGHT_KB")) p->xinflight = (unsigned int)std::max(0, atoi(v)) << 10;
        if (const char* v = getenv("SDNS_XFLUSH_CTAS")) p->xflush_ctas = std::max(1, atoi(v));
        if (const char* v = getenv("SDNS_XRATIO")) p->xratio = atof(v);
        p->kcopy = (xm && !strcmp(xm, "kcopy")) ? 1 : 0;    // send slots moved by a small copy kernel instead of the copy engines
        const char* env = getenv("SDNS_CHUNKS");
        p->nchunk = env ? atoi(env) : (p->xmode == 2 ? 6 : 4);
        if (p->nchunk < 1) p->nchunk = 1;
        if (p->nchunk > 16) p->nchunk = 16;
        if (!p->xmode) p->nchunk = 1;
        if (const char* kc = getenv("SDNS_KCOPY_CTAS")) p->kcopy_ctas = std::max(1, atoi(kc));
        if (const char* g = getenv("SDNS_GRAPH")) p->use_graph = atoi(g) != 0 && p->xmode >= 1;
        const char* sp = getenv("SDNS_SPLIT");
        p->nsplit = sp ? atoi(sp) : 1;      // measured: one copy stream per peer saturates the link (profiles/tools/p2p_copy_bench.py)
        if (p->nsplit < 1) p->nsplit = 1;
        if (p->nsplit > 4) p->nsplit = 4;
    }
    if (p->xmode == 1) {
        cudaError_t e1 = cudaSuccess;
        // a few copy streams shared by all destinations: one stream already drives the link at its rate
        // (profiles/tools/p2p_copy_bench.py); more only hide the per-copy start-up cost, and too many streams alias
        // onto the same hardware queues as the plan stream (CUDA_DEVICE_MAX_CONNECTIONS)
        const char* cse = getenv("SDNS_COPY_STREAMS");
        int ncs = cse ? atoi(cse) : 2;
        ncs = std::max(1, std::min(ncs, std::min(p->P - 1, 8)));
        for (int r = 0; r < ncs * p->nsplit && e1 == cudaSuccess; ++r) {
            cudaStream_t y = nullptr; cudaEvent_t e = nullptr;
            e1 = cudaStreamCreateWithFlags(&y, cudaStreamNonBlocking);
            p->ys.push_back(y);
            if (e1 == cudaSuccess) e1 = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            p->ev_y.push_back(e);
        }
        for (int c = 0; c < p->nchunk && e1 == cudaSuccess; ++c) {
            cudaEvent_t e = nullptr;
            e1 = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            p->ev_k.push_back(e);
        }
        if (e1 != cudaSuccess) { std::string m = std::string("copy streams/events: ") + cudaGetErrorString(e1); delete p; return fail(SDNS_ERR_CUDA, m); }
    }
    if (p->use_graph) {
        cudaError_t e1 = cudaStreamCreateWithFlags(&p->gstream, cudaStreamNonBlocking);
        if (e1 == cudaSuccess) e1 = cudaEventCreateWithFlags(&p->ev_gin, cudaEventDisableTiming);
        if (e1 == cudaSuccess) e1 = cudaEventCreateWithFlags(&p->ev_gout, cudaEventDisableTiming);
        if (e1 != cudaSuccess) { std::string m = std::string("graph stream/events: ") + cudaGetErrorString(e1); delete p; return fail(SDNS_ERR_CUDA, m); }
    }
    for (int i = 0; i < FAM_COUNT; ++i) { p->prof_ms[i] = 0; p->prof_bytes[i] = 0; p->prof_remote[i] = 0; p->prof_n[i] = 0; }
    build_spaces(p);
    for (int s = 0; s < 2; ++s)
        if (p->sp[s].M[0] < p->P) { delete p; return fail(SDNS_ERR_SIZE, "N[0] must be at least the number of ranks"); }
    for (int s = 0; s < 2; ++s) for (int i = 0; i < 3; ++i)
        if (!size_ok(p->sp[s].M[i])) {
            char b[128]; snprintf(b, sizeof b, "no compiled transform of length %d (have 2^k and 3*2^k, 8..3072, and 60, 90)", p->sp[s].M[i]);
            delete p; return fail(SDNS_ERR_SIZE, b);
        }
    {   // lengths with a factor 5 (60, 90) exist for the NS / VV Vortex path and the plain transforms only
        const bool vortex = cfg->solver != SDNS_MHD && cfg->convection == SDNS_CONV_VORTEX;
        for (int s = 0; s < 2 && !vortex; ++s) for (int i = 0; i < 3; ++i)
            if (p->sp[s].M[i] % 5 == 0) {
                delete p;
                return fail(SDNS_ERR_SIZE, "transform lengths 60 / 90 are compiled for the NS / VV Vortex path only");
            }
    }
    if (p->prec) fill_tables<double>(p); else fill_tables<float>(p);
    // scratch: A holds W0 (B0 out) and W2 (Z out); B holds W1 (B1 out) and W3 (F1 out)
    const int nz = cfg->solver == SDNS_MHD ? 9 : 6;    // widest field count through the pipeline
    // Every region is sized with the padded shares (M0p, N1p, the largest K1l of any rank) so that the workspace layout --
    // the offsets peers add to each other's base pointer -- is the same on all ranks, whatever their own shares are.
    size_t a = 0, b = 0, c = 0, sf = 0;
    for (int s = 0; s < 2; ++s) {
        const Space& q = p->sp[s];
        int k1lmax = 0;
        for (int r = 0; r < p->P; ++r) { const K1Own x = k1_own(p, q, r); k1lmax = std::max(k1lmax, x.a1 + x.nb); }
        size_t w0 = (size_t)6 * q.M0p * q.K1n * q.K2p;
        size_t w1 = (size_t)6 * q.M0p * q.M[1] * q.K2p;
        size_t w2 = (size_t)nz * q.M0p * q.M[1] * p->Nhp;
        size_t w3 = (size_t)nz * q.M[0] * p->N1p * p->Nhp;
        a = std::max(a, std::max(w0, w2));
        if (p->P == 1) b = std::max(b, std::max(w1, w3));      // single GPU: W3 reuses W1's buffer
        else { b = std::max(b, w1); c = std::max(c, w3); }      // multi GPU: peers write W3 while W1 is live
        if (p->xmode) {
            b = std::max(b, (size_t)6 * p->P * q.M0p * k1lmax * q.K2p);      // B0 send buffers: P slots of (6, M0p, K1l, K2p)
            sf = std::max(sf, (size_t)nz * p->P * p->N1p * q.M0p * p->Nhp);  // F1 send buffers: P slots of (nz, N1p, M0l, Nhp)
        }
    }
    p->bytes_A = align_up(a * p->cs, 256); p->bytes_B = align_up(b * p->cs, 256);
    p->bytes_C = align_up(c * p->cs, 256);
    p->red_blocks = SDNS_RED_BLOCKS;
    p->off_tab = 0;
    p->off_A = align_up(p->host_tables.size(), 256);
    p->off_B = p->off_A + p->bytes_A;
    p->off_C = p->P == 1 ? p->off_B : p->off_B + p->bytes_B;
    const bool needD = cfg->solver == SDNS_NS && (cfg->convection == SDNS_CONV_STANDARD || cfg->convection == SDNS_CONV_SKEWED);
    const bool needS = cfg->solver == SDNS_NS && cfg->convection == SDNS_CONV_SKEWED;
    p->bytes_D = needD ? align_up((size_t)3 * p->sp[1].M0p * p->sp[1].M[1] * p->Nhp * p->cs, 256) : 0;
    p->bytes_S = needS ? align_up((size_t)3 * p->N[0] * p->N1p * p->Nh * p->cs, 256) : 0;
    p->off_D = p->off_B + p->bytes_B + p->bytes_C;
    p->off_S = p->off_D + p->bytes_D;
    // inter-stage copy of the RK4 state in the k1-major work layout
    p->bytes_U = align_up((size_t)(cfg->solver == SDNS_MHD ? 6 : 3) * p->N[0] * p->N1p * p->Nh * p->cs, 256);
    p->off_U = p->off_S + p->bytes_S;
    p->bytes_SF = align_up(sf * p->cs, 256);
    p->off_SF = p->off_U + p->bytes_U;
    p->off_red = p->off_SF + p->bytes_SF;
    p->off_flags = p->off_red + align_up(sizeof(double) * std::max(p->red_blocks, 2 * SDNS_MAX_BINS), 256);
    p->ws_need = p->off_flags + 256;
    *out = p;
    return SDNS_OK;
}

extern "C" int sdns_plan_destroy(sdns_plan* p) {
    if (p) {
        for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
        if (p->tl_base) cudaEventDestroy(p->tl_base);
        if (p->gexec) cudaGraphExecDestroy(p->gexec);
        if (p->ev_gin) cudaEventDestroy(p->ev_gin);
        if (p->ev_gout) cudaEventDestroy(p->ev_gout);
        if (p->gstream) cudaStreamDestroy(p->gstream);
        for (cudaEvent_t e : p->ev_k) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : p->ev_y) if (e) cudaEventDestroy(e);
        for (cudaStream_t y : p->ys) if (y) cudaStreamDestroy(y);
        for (int r = 0; r < 8; ++r) if (p->peer_ws[r] && r != p->rank) cudaIpcCloseMemHandle(p->peer_ws[r]);
        if (p->own_ws && p->ws) cudaFree(p->ws);
    }
    delete p; return SDNS_OK;
}

extern "C" int sdns_workspace_bytes(const sdns_plan* p, size_t* bytes) {
    if (!p || !bytes) return fail(SDNS_ERR_ARG, "null argument");
    *bytes = p->ws_need; return SDNS_OK;
}

extern "C" int sdns_plan_set_workspace(sdns_plan* p, void* dptr, size_t bytes) {
    if (!p || !dptr) return fail(SDNS_ERR_ARG, "null argument");
    if (bytes < p->ws_need) return fail(SDNS_ERR_WORKSPACE, "workspace too small");
    if ((uintptr_t)dptr % 256) return fail(SDNS_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if (p->P > 1 && !p->own_ws)
        return fail(SDNS_ERR_STATE, "multi-GPU plans allocate their IPC-shared workspace with sdns_comm_alloc");
    p->ws = (char*)dptr; p->ws_bytes = bytes;
    p->peer_ws[p->rank] = p->ws;
    CUDA_TRY(cudaMemcpyAsync(p->ws + p->off_tab, p->host_tables.data(), p->host_tables.size(),
                             cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemsetAsync(p->ws + p->off_flags, 0, 256, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return SDNS_OK;
}

// flag-barrier timeout in clock64() ticks: ~10 s of SM clock.  The CPU emulation of the tests counts nanoseconds and
// runs every CUDA thread as an OS thread, so a rank can legitimately be tens of seconds behind on a loaded machine.
#ifdef SDNS_HOST_SHIM
#define SDNS_BARRIER_TIMEOUT 300000000000LL
#else
#define SDNS_BARRIER_TIMEOUT 20000000000LL
#endif

// ---- slab decomposition over NVLink peer memory (one process per GPU) -----------------------
// The all-to-all of mpi4py-fft's Transfer (MPI_Alltoallw; in-tree analogue spectralDNS3D_short.py:53,59)
// has no kernel of its own here: B0 and F1 store each output element directly into the owning rank's
// buffer.  That needs every rank's workspace mapped in every process: cudaMalloc + CUDA IPC.
extern "C" int sdns_comm_alloc(sdns_plan* p) {
    if (!p) return fail(SDNS_ERR_ARG, "null plan");
    if (p->ws) return fail(SDNS_ERR_STATE, "workspace already set");
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, p->ws_need));
    p->own_ws = true;
    int e = sdns_plan_set_workspace(p, d, p->ws_need);
    if (e) { cudaFree(d); p->own_ws = false; p->ws = nullptr; }
    return e;
}
extern "C" int sdns_comm_handle(sdns_plan* p, void* handle64) {
    if (!p || !handle64 || !p->own_ws) return fail(SDNS_ERR_STATE, "sdns_comm_handle: call sdns_comm_alloc first");
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, p->ws));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return SDNS_OK;
}
extern "C" int sdns_comm_open(sdns_plan* p, const void* handles, int nranks) {
    if (!p || !handles || nranks != p->P) return fail(SDNS_ERR_ARG, "sdns_comm_open: bad argument");
    if (!p->own_ws) return fail(SDNS_ERR_STATE, "sdns_comm_open: call sdns_comm_alloc first");
    for (int r = 0; r < nranks; ++r) {
        if (r == p->rank) { p->peer_ws[r] = p->ws; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * r, 64);
        void* d = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess));
        p->peer_ws[r] = (char*)d;
    }
    return SDNS_OK;
}

namespace sdns {
// ---------------------------------------------------------------------------------------
// Cross-GPU barrier over peer memory (one process per GPU, flags exchanged through CUDA IPC):
// thread r publishes this rank's epoch in rank r's flag array with a system-scope release, then
// waits for rank r's epoch in the local array.  Launched as <<<1, 32>>> between the pass that
// stores into the peers and the pass that reads what the peers stored.
// ---------------------------------------------------------------------------------------
struct BarrierArgs {
    unsigned int* peer_flags[8];     // peer_flags[r] = flag array living on rank r
    unsigned int* status;            // local: set to 1 on timeout
    int rank, nranks;
    unsigned int epoch;
    long long timeout_cycles;
};

__global__ void xbarrier_kernel(const BarrierArgs b) {
    const int r = threadIdx.x;
    if (r >= b.nranks) return;
    if (*reinterpret_cast<volatile unsigned int*>(b.status)) return;      // an earlier barrier timed out: the run is lost, do not wait again
    __threadfence_system();
    unsigned int* remote = b.peer_flags[r] + b.rank;
#ifndef SDNS_HOST_SHIM
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(b.epoch) : "memory");
#else
    __atomic_store_n(remote, b.epoch, __ATOMIC_RELEASE);
#endif
    const unsigned int* local = b.peer_flags[b.rank] + r;
    const long long t0 = clock64();
    unsigned int v;
    do {
#ifndef SDNS_HOST_SHIM
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
#else
        v = __atomic_load_n(local, __ATOMIC_ACQUIRE);
#endif
        if ((int)(v - b.epoch) >= 0) break;
        if (clock64() - t0 > b.timeout_cycles) { *b.status = 1u; break; }
    } while (true);
    __threadfence_system();
}

// Barrier for captured (graph) steps: identical handshake, but the epoch comes from a counter in device memory that
// the kernel itself advances (kernel arguments are frozen in a graph), and it uses its own flag words so that eager
// and captured barriers never see each other's epochs.
struct BarrierDevArgs {
    unsigned int* peer_flags[8];
    unsigned int* status;
    unsigned int* counter;
    int rank, nranks;
    long long timeout_cycles;
};
__global__ void xbarrier_dev_kernel(const BarrierDevArgs b) {
    SDNS_STATIC_SMEM(unsigned int, sh, 1);
    if (threadIdx.x == 0) sh[0] = ++(*b.counter);
    __syncthreads();
    const unsigned int epoch = sh[0];
    const int r = threadIdx.x;
    if (r >= b.nranks) return;
    if (*reinterpret_cast<volatile unsigned int*>(b.status)) return;
    __threadfence_system();
    unsigned int* remote = b.peer_flags[r] + b.rank;
#ifndef SDNS_HOST_SHIM
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
#else
    __atomic_store_n(remote, epoch, __ATOMIC_RELEASE);
#endif
    const unsigned int* local = b.peer_flags[b.rank] + r;
    const long long t0 = clock64();
    unsigned int v;
    do {
#ifndef SDNS_HOST_SHIM
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
#else
        v = __atomic_load_n(local, __ATOMIC_ACQUIRE);
#endif
        if ((int)(v - epoch) >= 0) break;
        if (clock64() - t0 > b.timeout_cycles) { *b.status = 1u; break; }
    } while (true);
    __threadfence_system();
}

}  // namespace sdns

static cudaEvent_t get_event(sdns_plan* p);

static int xbarrier(sdns_plan* p) {
    if (p->P == 1) return SDNS_OK;
    for (int r = 0; r < p->P; ++r) if (!p->peer_ws[r]) return fail(SDNS_ERR_STATE, "peers not opened (sdns_comm_open)");
    if (p->capturing) {
        // flag words 16..23 and the counter at word 40 of the 64-word flag block belong to the captured barriers
        BarrierDevArgs d;
        for (int r = 0; r < 8; ++r) d.peer_flags[r] = r < p->P ? reinterpret_cast<unsigned int*>(p->peer_ws[r] + p->off_flags) + 16 : nullptr;
        d.status = reinterpret_cast<unsigned int*>(p->ws + p->off_flags) + 32;
        d.counter = reinterpret_cast<unsigned int*>(p->ws + p->off_flags) + 40;
        d.rank = p->rank; d.nranks = p->P; d.timeout_cycles = SDNS_BARRIER_TIMEOUT;
        SDNS_LAUNCH(xbarrier_dev_kernel, 1, 32, 0, p->stream)(d);
        p->launches++;
        CUDA_TRY(cudaGetLastError());
        return SDNS_OK;
    }
    BarrierArgs b;
    for (int r = 0; r < 8; ++r) b.peer_flags[r] = r < p->P ? reinterpret_cast<unsigned int*>(p->peer_ws[r] + p->off_flags) : nullptr;
    b.status = reinterpret_cast<unsigned int*>(p->ws + p->off_flags) + 32;
    b.rank = p->rank; b.nranks = p->P; b.epoch = ++p->epoch;
    b.timeout_cycles = SDNS_BARRIER_TIMEOUT;      // ~10 s: a missing peer must not hang the GPU
    sdns_plan::Rec r; r.fam = -1; r.bytes = 0; r.remote = 0;
    if (p->tl_on) { r.a = get_event(p); cudaEventRecord(r.a, p->stream); }
    SDNS_LAUNCH(xbarrier_kernel, 1, 32, 0, p->stream)(b);
    if (p->tl_on) { r.b = get_event(p); cudaEventRecord(r.b, p->stream); p->brecs.push_back(r); }
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}

extern "C" int sdns_comm_status(sdns_plan* p, int* timed_out) {
    if (!p || !timed_out) return fail(SDNS_ERR_ARG, "null argument");
    unsigned int v = 0;
    if (p->ws) {
        CUDA_TRY(cudaMemcpyAsync(&v, p->ws + p->off_flags + 32 * sizeof(unsigned int), sizeof v, cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
    }
    *timed_out = (int)v;
    return SDNS_OK;
}

// A barrier timeout means a peer was slow, hung or dead and every later pass ran on partially exchanged data: the
// fault is sticky (the barriers stop waiting) and fatal at every point where results leave the device.
static int check_comm(sdns_plan* p) {
    if (p->P == 1 || !p->ws) return SDNS_OK;
    unsigned int v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, p->ws + p->off_flags + 32 * sizeof(unsigned int), sizeof v, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (v) return fail(SDNS_ERR_STATE, "cross-GPU barrier timed out (a peer rank is slow, hung or gone): results since then are invalid");
    return SDNS_OK;
}

extern "C" int sdns_plan_set_stream(sdns_plan* p, void* s) {
    if (!p) return fail(SDNS_ERR_ARG, "null plan");
    p->stream = (cudaStream_t)s; return SDNS_OK;
}
extern "C" int sdns_sync(sdns_plan* p) {
    if (!p) return fail(SDNS_ERR_ARG, "null plan");
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return check_comm(p);
}
extern "C" int sdns_local_shapes(const sdns_plan* p, int32_t sp[3], int32_t ph[3], int32_t pd[3]) {
    if (!p) return fail(SDNS_ERR_ARG, "null plan");
    sp[0] = p->N[0]; sp[1] = p->N1l; sp[2] = p->Nh;
    for (int i = 0; i < 3; ++i) { ph[i] = p->sp[0].M[i]; pd[i] = p->sp[1].M[i]; }
    ph[0] = p->sp[0].M0l; pd[0] = p->sp[1].M0l;
    return SDNS_OK;
}
extern "C" int sdns_k1_layout(const sdns_plan* p, int32_t* first, int32_t* step) {
    if (!p || !first || !step) return fail(SDNS_ERR_ARG, "null argument");
    *first = p->k1cyc ? p->rank : p->s1;
    *step = p->k1cyc ? p->P : 1;
    return SDNS_OK;
}
extern "C" int sdns_launch_count(const sdns_plan* p, long long* c) {
    if (!p || !c) return fail(SDNS_ERR_ARG, "null argument");
    *c = p->launches; return SDNS_OK;
}

static int need_ws(sdns_plan* p) {
    if (!p) return fail(SDNS_ERR_ARG, "null plan");
    if (!p->ws) return fail(SDNS_ERR_WORKSPACE, "call sdns_plan_set_workspace first");
    return SDNS_OK;
}

static cudaEvent_t get_event(sdns_plan* p) {
    if (p->ev_used == p->ev_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); p->ev_pool.push_back(e); }
    return p->ev_pool[p->ev_used++];
}

// bytes = algorithmic HBM bytes of this launch: every input element read once + every output
// element written once at the pass's actual (pruned / padded) sizes
static int do_launch(sdns_plan* p, cudaStream_t st, int fam, int n, const void* args, double bytes = 0, double remote = 0) {
    sdns_plan::Rec r; r.fam = fam; r.bytes = bytes; r.remote = remote;
    if (p->prof) { r.a = get_event(p); cudaEventRecord(r.a, st); }
    int e = g_launch[fam][p->prec](n, args, st);
    if (p->prof) { r.b = get_event(p); cudaEventRecord(r.b, st); p->recs.push_back(r); }
    p->launches++;
    if (e == -1001) return fail(SDNS_ERR_SIZE, "array too large for one pass: transform length x line stride must stay below 2^31 elements per field (use more GPUs)");
    if (e == -1000) { char b[96]; snprintf(b, sizeof b, "no kernel for length %d (family %d)", n, fam); return fail(SDNS_ERR_SIZE, b); }
    if (e != 0) { char b[160]; snprintf(b, sizeof b, "kernel launch (family %d, n=%d): %s", fam, n, cudaGetErrorString((cudaError_t)e)); return fail(SDNS_ERR_CUDA, b); }
    return SDNS_OK;
}

// ---- transfer role (xmode 2, xfer.cuh) ---------------------------------------------------
namespace sdns {
__global__ void __launch_bounds__(128) xfer_kernel(const XferOnlyArgs a) {
    SDNS_DYN_SMEM(smraw);
    xfer_role(a.x, smraw, blockIdx.x, gridDim.x);
}
}  // namespace sdns

// One 2-D copy per destination, all of the same shape, queued until pass launches of the plan stream carry it.
static int push_xfer(sdns_plan* p, int ndest, const void* const* src, void* const* dst, size_t spitch, size_t dpitch,
                     size_t width, size_t height) {
    if (!ndest || !width || !height) return SDNS_OK;
    if (ndest > SDNS_XD_MAX || width >= (1ull << 32) || height >= (1ull << 32)) return fail(SDNS_ERR_ARG, "push_xfer: shape");
    XferBatch b; memset(&b, 0, sizeof b);
    for (int d = 0; d < ndest; ++d) {
        b.src[d] = (const char*)src[d]; b.dst[d] = (char*)dst[d];
        if ((((uintptr_t)src[d]) | ((uintptr_t)dst[d])) & 15) return fail(SDNS_ERR_STATE, "push_xfer: send slots must be 16-byte aligned");
    }
    if ((spitch | dpitch | width) & 15) return fail(SDNS_ERR_STATE, "push_xfer: rows must be multiples of 16 bytes");
    b.spitch = spitch; b.dpitch = dpitch; b.width = (unsigned int)width; b.row0 = 0; b.nrows = (unsigned int)height; b.ndest = ndest;
    p->pend.push_back(b);
    p->xfer_bytes += (double)ndest * width * height;
    return SDNS_OK;
}
// the launch about to be made moves `hbm_bytes` through HBM: let it carry xratio times that over NVLink
static void attach_xfer(sdns_plan* p, XferArgs& x, double hbm_bytes) {
    x.nctas = 0; x.nbatch = 0; x.tma = p->xtma;
    if (p->xmode != 2 || p->pend.empty()) return;
    double budget = hbm_bytes * p->xratio;
    while (!p->pend.empty() && x.nbatch < SDNS_XB_MAX && budget > 0) {
        XferBatch& j = p->pend.front();
        const double per_row = (double)j.width * j.ndest;
        unsigned int take = (unsigned int)std::min<double>(j.nrows, std::max(1.0, floor(budget / per_row)));
        XferBatch t = j; t.nrows = take;
        x.b[x.nbatch++] = t;
        j.row0 += take; j.nrows -= take; budget -= take * per_row;
        if (!j.nrows) p->pend.pop_front();
    }
    if (x.nbatch) { x.nctas = p->xctas; x.inflight = p->xinflight; }
}
// everything still pending, as a launch of its own (before the barrier in front of the pass that needs the data)
static int flush_xfer(sdns_plan* p) {
    while (!p->pend.empty()) {
        XferOnlyArgs a; memset(&a, 0, sizeof a);
        a.x.tma = p->xtma;
        while (!p->pend.empty() && a.x.nbatch < SDNS_XB_MAX) { a.x.b[a.x.nbatch++] = p->pend.front(); p->pend.pop_front(); }
        a.x.nctas = p->xflush_ctas;
#ifdef SDNS_HOST_SHIM
        const size_t smem = 4096;
#else
        const size_t smem = 96 * 1024;
        static bool once = false;
        if (!once) { CUDA_TRY(cudaFuncSetAttribute(xfer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once = true; }
#endif
        xfer_prepare(a.x, smem, 128);
        sdns_plan::Rec r; r.fam = -2; r.bytes = 0; r.remote = 0;
        if (p->tl_on) { r.a = get_event(p); cudaEventRecord(r.a, p->stream); }
        SDNS_LAUNCH(xfer_kernel, a.x.nctas, 128, smem, p->stream)(a);
        if (p->tl_on) { r.b = get_event(p); cudaEventRecord(r.b, p->stream); p->brecs.push_back(r); }
        p->launches++; p->xfer_flushes++;
        CUDA_TRY(cudaGetLastError());
    }
    return SDNS_OK;
}

// ---- typed pipeline --------------------------------------------------------------------
struct Rng { int a, b; };     // half-open index range of a chunked launch; a < 0: the whole axis

template <typename T>
struct Pipe {
    typedef typename C2<T>::type V;
    sdns_plan* p;
    const Space& q;
    V* A; V* B; V* C;
    cudaStream_t st;            // stream the next pass is launched on
    int zcap;                   // resident CTAs per SM for the persistent z kernel (0 = all)
    int xcap;                   // CTAs per SM of the passes that store into the peers (0 = one CTA per tile)
    Pipe(sdns_plan* p_, int space) : p(p_), q(p_->sp[space]), st(p_->stream), zcap(0), xcap(0) {
        A = reinterpret_cast<V*>(p->ws + p->off_A); B = reinterpret_cast<V*>(p->ws + p->off_B);
        C = reinterpret_cast<V*>(p->ws + p->off_C);
    }
    const V* tw(int n) const { return reinterpret_cast<const V*>(p->ws + p->off_tab + p->tw_off.at(n)); }
    void base(StridedArgs<T>& a) const {
        memset(&a, 0, sizeof a);
        a.kx = reinterpret_cast<const T*>(p->ws + p->kx_off);
        a.ky = reinterpret_cast<const T*>(p->ws + p->ky_off);
        a.kz = reinterpret_cast<const T*>(p->ws + p->kz_off);
        a.N0 = p->N[0]; a.N1 = p->N[1]; a.N2 = p->N[2];
        a.mask_nyquist = p->cfg.mask_nyquist;
        a.scale = (T)1;
        a.st_fs = dense_fs();
        a.k1_off = p->k1cyc ? p->rank : p->s1; a.k1_mul = p->k1cyc ? p->P : 1;
        a.uh_ls = (long long)p->N1l * p->Nh; a.uh_os = p->Nh;          // reference layout (N0, N1l, Nh)
        a.t_ls = p->Nh; a.t_os = (long long)p->N[0] * p->Nh;           // work layout (N1l, N0, Nh)
    }
    long long dense_fs() const { return (long long)p->N[0] * p->N1l * p->Nh; }
    // destinations of a pass that feeds a global transpose.  Direct: every rank's final array (peer memory).
    // Staged (send != null): this rank's part goes to its final array, rank r's part to slot r of the local send
    // buffer in the layout (fs2, ls2, c1off2) from which the copy engines move it.
    void peers(StridedArgs<T>& a, size_t off, int chunk, V* send = nullptr, long long slot = 0,
               long long fs2 = 0, long long ls2 = 0, long long c1off2 = 0) const {
        a.self = -1; a.out_fs2 = a.out_fs; a.out_ls2 = a.out_ls; a.c1_out_off2 = a.c1_out_off;
        if (p->P == 1) return;
        a.xchunk = chunk;
        for (int r = 0; r < p->P; ++r) a.peer_out[r] = reinterpret_cast<V*>(p->peer_ws[r] + off);
        if (send) {
            for (int r = 0; r < p->P; ++r) if (r != p->rank) a.peer_out[r] = send + r * slot;
            a.self = p->rank; a.out_fs2 = fs2; a.out_ls2 = ls2; a.c1_out_off2 = c1off2;
        }
    }
    static Rng whole(Rng r, int n) { if (r.a < 0) { r.a = 0; r.b = n; } return r; }

    // B0: local dense spectral (nf, N0, N1l, Nh) -> W0 (nfo, M0l, K1n, K2p) of the rank owning each x0.
    // k2: the k2 columns of this launch.
    long long b0_slot() const { return (long long)6 * q.M0p * q.K1l * q.K2p; }
    long long f1_slot(int nf) const { return (long long)nf * p->N1p * q.M0l * p->Nhp; }
    V* send_f1() const { return reinterpret_cast<V*>(p->ws + p->off_SF); }

    int b0(int fam, const V* in, int nf, int comp = 0, bool work_layout = false, Rng k2 = Rng{-1, 0},
           Rng k1 = Rng{-1, 0}, bool staged = false) {
        k2 = whole(k2, q.K2n);
        k1 = whole(k1, q.K1l);                          // compact local axis-1 modes of this launch
        StridedArgs<T> a; base(a);
        a.in = in; a.out = A; a.comp = comp;
        a.in_fs = dense_fs();
        if (work_layout) { a.in_ls = p->Nh; a.in_os = (long long)p->N[0] * p->Nh; }
        else { a.in_ls = (long long)p->N1l * p->Nh; a.in_os = p->Nh; }
        a.cw = k2.b - k2.a; a.c2_off = k2.a; a.c1_off = k1.a; a.ncols = (long long)(k1.b - k1.a) * a.cw;
        a.col_nlo = q.lcol_nlo; a.col_gap = q.lcol_gap;
        a.imap = q.bmap[0]; a.omap = all_map(q.M[0]);
        a.out_fs = (long long)q.M0p * q.K1n * q.K2p; a.out_ls = (long long)q.K1n * q.K2p; a.out_os = q.K2p;
        a.c1_out_off = q.c1off;
        if (staged) peers(a, p->off_A, q.M0p, B, b0_slot(), (long long)q.M0p * q.K1l * q.K2p, (long long)q.K1l * q.K2p, 0);
        else peers(a, p->off_A, q.M0p);
        if (p->P > 1 && q.M[0] % p->P) a.otab = reinterpret_cast<const int*>(p->ws + p->off_tab + q.xtab_off);
        a.grid_cap = xcap;
        a.tw = tw(q.M[0]); a.nfields = nf;
        const int nfo = (fam == FAM_PLAIN_BWD) ? nf : 6;
        const double cols = (double)(k1.b - k1.a) * a.cw;
        const double bytes = (nf * cols * (q.bmap[0].nlo + q.bmap[0].nhi) + nfo * cols * q.M[0]) * p->cs;
        if (a.ncols == 0) return SDNS_OK;              // this rank owns no mode that survives the truncation
        const double remote = staged ? 0 : nfo * cols * q.M[0] * p->cs * (p->P - 1) / p->P;   // stored into peers over NVLink
        attach_xfer(p, a.x, bytes);
        return do_launch(p, st, fam, q.M[0], &a, bytes, remote);
    }
    // B1: A (W0) -> B as W1 (nf, M0l, M1, K2p)
    int b1(int nf, Rng k2 = Rng{-1, 0}) {
        k2 = whole(k2, q.K2n);
        StridedArgs<T> a; base(a);
        a.in = A; a.out = B;
        a.in_fs = (long long)q.M0p * q.K1n * q.K2p; a.in_ls = q.K2p; a.in_os = (long long)q.K1n * q.K2p;
        a.cw = k2.b - k2.a; a.c2_off = k2.a; a.ncols = (long long)q.M0l * a.cw;
        a.col_nlo = q.M0l; a.col_gap = 0;
        a.imap = q.bmap[1];
        if (p->k1cyc) {
            a.itab = reinterpret_cast<const int*>(p->ws + p->off_tab + q.itab_off);
            a.cycP = p->P; a.cyc_dm = q.cyc_dm;
            for (int r = 0; r < 8; ++r) { a.cyc_first[r] = q.cyc_first[r]; a.cyc_hi[r] = q.cyc_hi[r]; }
        }
        a.omap = all_map(q.M[1]);
        a.out_fs = (long long)q.M0l * q.M[1] * q.K2p; a.out_ls = q.K2p; a.out_os = (long long)q.M[1] * q.K2p;
        a.tw = tw(q.M[1]); a.nfields = nf;
        const double bytes = (double)nf * q.M0l * a.cw * ((double)q.K1n + q.M[1]) * p->cs;
        attach_xfer(p, a.x, bytes);
        return do_launch(p, st, FAM_PLAIN_BWD, q.M[1], &a, bytes);
    }
    // Z: B (W1) -> A as W2 (nfo, M0l, M1, Nhp)   [fused], or to/from user real arrays [plain].
    // x0: the local x0 planes of this launch.
    int z(int fam, const void* in, void* out, int nf, bool in_is_W1, bool out_is_W2, Rng x0 = Rng{-1, 0}) {
        x0 = whole(x0, q.M0l);
        ZArgs<T> a; memset(&a, 0, sizeof a);
        const long long plane = (long long)q.M0l * q.M[1];
        a.in_ls = in_is_W1 ? q.K2p : q.M[2]; a.in_fs = plane * a.in_ls;
        a.out_ls = out_is_W2 ? p->Nhp : q.M[2]; a.out_fs = plane * a.out_ls;
        const long long l0 = (long long)x0.a * q.M[1];
        a.in = in_is_W1 ? (const void*)(reinterpret_cast<const V*>(in) + l0 * a.in_ls)
                        : (const void*)(reinterpret_cast<const T*>(in) + l0 * a.in_ls);
        a.out = out_is_W2 ? (void*)(reinterpret_cast<V*>(out) + l0 * a.out_ls)
                          : (void*)(reinterpret_cast<T*>(out) + l0 * a.out_ls);
        a.nlines = (long long)(x0.b - x0.a) * q.M[1]; a.nin_keep = q.K2n; a.nout_keep = p->Nh; a.nf = nf;
        a.tw = tw(q.M[2]);
        a.scale = (fam == FAM_Z_C2R) ? (T)1 : (T)q.scale;
        a.grid_cap = zcap;
        const int nin = (fam == FAM_Z_CROSS || fam == FAM_Z_MHD || fam == FAM_Z_DOT) ? 6 : (fam == FAM_Z_UU ? 3 : nf);
        const int nout = fam == FAM_Z_CROSS ? 3 : (fam == FAM_Z_MHD ? 9 : (fam == FAM_Z_DOT ? 1 : (fam == FAM_Z_UU ? 6 : nf)));
        const double bin = in_is_W1 ? (double)q.K2n * p->cs : (double)q.M[2] * p->rs;
        const double bout = out_is_W2 ? (double)p->Nh * p->cs : (double)q.M[2] * p->rs;
        if (a.nlines == 0) return SDNS_OK;
        attach_xfer(p, a.x, (double)a.nlines * (nin * bin + nout * bout));
        return do_launch(p, st, fam, q.M[2], &a, (double)a.nlines * (nin * bin + nout * bout));
    }
    // F1: A (W2) -> W3 (nf, N1l, M0, Nhp) of the rank owning each k1.  x0 is the second-fastest axis of
    // W3 on purpose: F1 *stores* with the large stride (M0*Nhp), F0 *loads* its axis-0 lines with a
    // stride of one k2 row -- loads stall warps on TLB/DRAM-page misses, stores do not.
    int f1(int nf, const V* src = nullptr, Rng x0 = Rng{-1, 0}, bool staged = false) {
        x0 = whole(x0, q.M0l);
        StridedArgs<T> a; base(a);
        a.in = src ? src : A; a.out = C;
        a.in_fs = (long long)q.M0l * q.M[1] * p->Nhp; a.in_ls = p->Nhp; a.in_os = (long long)q.M[1] * p->Nhp;
        a.cw = p->Nh; a.c1_off = x0.a; a.ncols = (long long)(x0.b - x0.a) * p->Nh;
        a.col_nlo = q.M0l; a.col_gap = 0;
        a.imap = all_map(q.M[1]); a.omap = q.fmap[1];
        a.out_fs = (long long)p->N1p * q.M[0] * p->Nhp; a.out_ls = (long long)q.M[0] * p->Nhp; a.out_os = p->Nhp;
        a.c1_out_off = q.s0;
        if (staged) peers(a, p->off_C, p->N1p, send_f1(), f1_slot(nf), (long long)p->N1p * q.M0l * p->Nhp, (long long)q.M0l * p->Nhp, 0);
        else peers(a, p->off_C, p->N1p);
        if (p->k1cyc) { a.otab = reinterpret_cast<const int*>(p->ws + p->off_tab + q.otab_off); a.cycP = p->P; }
        else if (p->uneven1 && p->P > 1) a.otab = reinterpret_cast<const int*>(p->ws + p->off_tab + q.otab_off);
        a.grid_cap = xcap;
        a.tw = tw(q.M[1]); a.nfields = nf;
        const double bytes = (double)nf * (x0.b - x0.a) * p->Nh * ((double)q.M[1] + p->N[1]) * p->cs;
        const double remote = staged ? 0 : (double)nf * (x0.b - x0.a) * p->Nh * p->N[1] * p->cs * (p->P - 1) / p->P;
        if (a.ncols == 0) return SDNS_OK;
        attach_xfer(p, a.x, bytes);
        return do_launch(p, st, FAM_PLAIN_FWD, q.M[1], &a, bytes, remote);
    }
    // F0 geometry: W3 -> local dense spectral, k2 columns [k2.a, k2.b)
    void f0_geom(StridedArgs<T>& a, int nf, Rng k2 = Rng{-1, 0}, Rng k1 = Rng{-1, 0}) {
        k2 = whole(k2, p->Nh);
        k1 = whole(k1, p->N1l);
        base(a);
        a.in = C;
        a.in_fs = (long long)p->N1p * q.M[0] * p->Nhp; a.in_ls = p->Nhp; a.in_os = (long long)q.M[0] * p->Nhp;
        a.cw = k2.b - k2.a; a.c2_off = k2.a; a.c1_off = k1.a; a.ncols = (long long)(k1.b - k1.a) * a.cw;
        a.col_nlo = p->N1l; a.col_gap = 0;
        a.imap = all_map(q.M[0]); a.omap = q.fmap[0];
        a.out_fs = dense_fs(); a.out_ls = (long long)p->N1l * p->Nh; a.out_os = p->Nh;
        a.tw = tw(q.M[0]); a.nfields = nf;
    }
};

template <typename T>
static int backward_t(sdns_plan* p, int space, int nc, const void* in, void* out) {
    typedef typename C2<T>::type V;
    Pipe<T> P(p, space);
    for (int c0 = 0; c0 < nc; c0 += 6) {
        int nf = std::min(6, nc - c0);
        const V* src = reinterpret_cast<const V*>(in) + (long long)c0 * P.dense_fs();
        T* dst = reinterpret_cast<T*>(out) + (long long)c0 * P.q.M0l * P.q.M[1] * P.q.M[2];
        int e;
        if ((e = P.b0(FAM_PLAIN_BWD, src, nf))) return e;
        if ((e = xbarrier(p))) return e;
        if ((e = P.b1(nf))) return e;
        // every rank has finished READING W0 before any rank's next operation stores into it again (a faster
        // peer could otherwise start its next B0 while this rank is still in B1)
        if ((e = xbarrier(p))) return e;
        if ((e = P.z(FAM_Z_C2R, P.B, dst, nf, true, false))) return e;
    }
    return SDNS_OK;
}

template <typename T>
static int forward_t(sdns_plan* p, int space, int nc, const void* in, void* out) {
    typedef typename C2<T>::type V;
    Pipe<T> P(p, space);
    const int chunk = p->cfg.solver == SDNS_MHD ? 9 : 6;
    for (int c0 = 0; c0 < nc; c0 += chunk) {
        int nf = std::min(chunk, nc - c0);
        const T* src = reinterpret_cast<const T*>(in) + (long long)c0 * P.q.M0l * P.q.M[1] * P.q.M[2];
        V* dst = reinterpret_cast<V*>(out) + (long long)c0 * P.dense_fs();
        int e;
        // no rank may still be reading W3 (the F0 of the previous operation) when the first F1 stores into it
        if ((e = xbarrier(p))) return e;
        if ((e = P.z(FAM_Z_R2C, src, P.A, nf, false, true))) return e;
        if ((e = P.f1(nf))) return e;
        if ((e = xbarrier(p))) return e;
        StridedArgs<T> a; P.f0_geom(a, nf);
        a.out = dst;
        const double bytes = (double)nf * p->N1l * p->Nh * ((double)P.q.M[0] + p->N[0]) * p->cs;
        if ((e = do_launch(p, p->stream, FAM_PLAIN_FWD, P.q.M[0], &a, bytes))) return e;
    }
    return SDNS_OK;
}

struct StageOut {
    int out_mode; void* rhs; void* u0; void* u1; void* u2; void* p_hat; const void* source;
    double adt, bdt; int rk;
    bool in_work_layout;      // u_hat is the library's k1-major inter-stage copy
    const void* next_u;       // input of the right-hand side that follows immediately (RK4 stages 0-2: the
                              // inter-stage state this stage writes), or NULL: lets the multi-GPU pipeline start
                              // that stage's B0 chunks and their copies between this stage's F0 chunks
};

// the F0 launch (final forward pass + epilogue) over the k2 columns `k2` of the nprod product fields
template <typename T>
static int launch_f0(sdns_plan* p, Pipe<T>& P, const void* u_hat, double nu, double eta, const StageOut& so,
                     int nprod, bool divform, T f0scale, Rng k2, Rng k1 = Rng{-1, 0}) {
    typedef typename C2<T>::type V;
    const int solver = p->cfg.solver, conv = p->cfg.convection;
    k2 = Pipe<T>::whole(k2, p->Nh);
    k1 = Pipe<T>::whole(k1, p->N1l);
    if (k2.b <= k2.a || k1.b <= k1.a) return SDNS_OK;
    StridedArgs<T> a; P.f0_geom(a, nprod, k2, k1);
    a.out_mode = so.out_mode;
    a.u_hat = reinterpret_cast<const V*>(u_hat);
    a.rhs = reinterpret_cast<V*>(so.rhs);
    a.u0 = reinterpret_cast<V*>(so.u0); a.u1 = reinterpret_cast<V*>(so.u1); a.u2 = reinterpret_cast<V*>(so.u2);
    a.source = reinterpret_cast<const V*>(so.source);
    a.p_hat = reinterpret_cast<V*>(so.p_hat);
    a.nu = (T)nu; a.eta = (T)eta; a.adt = (T)so.adt; a.bdt = (T)so.bdt; a.rk = so.rk;
    a.scale = f0scale;
    if (so.in_work_layout) { a.uh_ls = a.t_ls; a.uh_os = a.t_os; }
    if (divform) {
        a.cfac = conv == SDNS_CONV_SKEWED ? (T)-0.5 : (T)-1;
        a.addin = conv == SDNS_CONV_SKEWED ? reinterpret_cast<const V*>(p->ws + p->off_S) : nullptr;
    }
    const int fam = divform ? FAM_NSDIV_F0 : (solver == SDNS_NS ? FAM_NS_F0 : (solver == SDNS_VV ? FAM_VV_F0 : FAM_MHD_F0));
    // epilogue traffic: read the product fields; state reads/writes of the stage update
    const int ns = solver == SDNS_MHD ? 6 : 3;
    double stt;   // state arrays touched, in units of one ns-component spectral vector
    if (so.out_mode == OUT_RHS) stt = 2;                               // read u_hat, write rhs
    else if (so.out_mode == OUT_CONV) stt = 1;
    else stt = so.rk == 0 ? 1 + 3 : (so.rk < 3 ? 3 + 2 : 2 + 1);      // see passes.cuh RK4 stage
    const double w = (double)(k2.b - k2.a) * (k1.b - k1.a);       // columns of this launch
    const double dense = (double)p->N[0] * w * p->cs;
    const double bytes = (double)nprod * w * P.q.M[0] * p->cs + stt * ns * dense
                         + (so.source ? ns * dense : 0) + (so.p_hat ? dense : 0);
    attach_xfer(p, a.x, bytes);
    return do_launch(p, P.st, fam, P.q.M[0], &a, bytes);
}

namespace sdns {
// SDNS_EXCHANGE=kcopy: the strided copy of a send slot into the peer's array as a kernel -- 16-byte loads from local
// HBM, 16-byte stores over NVLink, a few CTAs only (grid-stride) so that it runs beside the passes of the plan
// stream.  Same rows / pitches as the copy-engine path; an experiment to separate copy-engine behaviour from the
// schedule.
__global__ void slot_copy_kernel(uint4* __restrict__ dst, size_t dpitch16, const uint4* __restrict__ src, size_t spitch16,
                                 size_t width16, size_t height) {
    const size_t total = width16 * height;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / width16, c = i - r * width16;
        dst[r * dpitch16 + c] = src[r * spitch16 + c];
    }
}
}  // namespace sdns

// One 2-D copy per destination rank on that rank's copy stream, after `after` has fired
static int copy_rows(sdns_plan* p, int r, cudaEvent_t after, void* dst, size_t dpitch, const void* src, size_t spitch,
                     size_t width, size_t height) {
    if (!width || !height) return SDNS_OK;
    for (int i = 0; i < p->nsplit; ++i) {
        const size_t h0 = height * i / p->nsplit, h1 = height * (i + 1) / p->nsplit;
        if (h1 == h0) continue;
        // destinations are visited nearest-neighbour first (rank+1, rank+2, ...) so that no rank is every sender's
        // first target; destination k of that order uses stream k mod (number of streams)
        const int ncs = (int)p->ys.size() / p->nsplit;
        const int si = (((r - p->rank - 1 + p->P) % p->P) % ncs) * p->nsplit + i;
        cudaStream_t y = p->ys[si];
        CUDA_TRY(cudaStreamWaitEvent(y, after, 0));
        sdns_plan::CRec cr; cr.s = si; cr.bytes = (double)width * (h1 - h0);
        if (p->prof) { cr.a = get_event(p); cudaEventRecord(cr.a, y); }
        if (p->kcopy && width % 16 == 0 && dpitch % 16 == 0 && spitch % 16 == 0) {
            SDNS_LAUNCH(slot_copy_kernel, p->kcopy_ctas, 256, 0, y)(reinterpret_cast<uint4*>((char*)dst + h0 * dpitch), dpitch / 16,
                                                                   reinterpret_cast<const uint4*>((const char*)src + h0 * spitch), spitch / 16,
                                                                   width / 16, h1 - h0);
            CUDA_TRY(cudaGetLastError());
        } else
        CUDA_TRY(cudaMemcpy2DAsync((char*)dst + h0 * dpitch, dpitch, (const char*)src + h0 * spitch, spitch, width, h1 - h0,
                                   cudaMemcpyDeviceToDevice, y));
        if (p->prof) { cr.b = get_event(p); cudaEventRecord(cr.b, y); p->crecs.push_back(cr); }
    }
    return SDNS_OK;
}
// the plan stream waits until every copy stream has drained, then the cross-GPU barrier
static int join_copies(sdns_plan* p) {
    if (p->xmode == 2) { int e = flush_xfer(p); if (e) return e; return xbarrier(p); }
    for (size_t si = 0; si < p->ys.size(); ++si) {
        if (!p->ys[si]) continue;
        CUDA_TRY(cudaEventRecord(p->ev_y[si], p->ys[si]));
        CUDA_TRY(cudaStreamWaitEvent(p->stream, p->ev_y[si], 0));
    }
    return xbarrier(p);
}

// Chunk boundary c (0..nc) of an axis of n entries.  The chunks shrink towards the end (weights 3,..,3,2,1): the
// copy of the last chunk is the only one no later pass hides.
static int chunk_bound(int n, int nc, int c) {
    auto w = [&](int i) { return std::min(nc - i, 3); };
    int tot = 0, cum = 0;
    for (int i = 0; i < nc; ++i) { tot += w(i); if (i < c) cum += w(i); }
    return (int)(((long long)n * cum + tot / 2) / tot);
}

// Chunk c of nc over this rank's KEPT axis-1 modes (compact local indices, what B0 transforms), and the one or two
// memory ranges of the spectral arrays that hold them (what F0 must have written before).  The modes the 2/3 rule
// truncates belong to no chunk: F0 handles them last (k1_truncated), underneath the last B0 copies.
static void k1_chunk(const sdns_plan* p, const Space& q, int c, int nc, Rng* kept, Rng mem[2]) {
    kept->a = chunk_bound(q.K1l, nc, c); kept->b = chunk_bound(q.K1l, nc, c + 1);
    mem[0] = Rng{std::min(kept->a, q.lcol_nlo), std::min(kept->b, q.lcol_nlo)};                     // low run: memory = compact
    mem[1] = Rng{std::max(kept->a, q.lcol_nlo) + q.lcol_gap, std::max(kept->b, q.lcol_nlo) + q.lcol_gap};   // high run
    (void)p;
}
static Rng k1_truncated(const sdns_plan* p, const Space& q) {
    return q.K1l > q.lcol_nlo ? Rng{q.lcol_nlo, q.lcol_nlo + q.lcol_gap} : Rng{q.lcol_nlo, p->N1l};
}

// B0 chunk c of the right-hand side with input u: the pass (own part straight into W0, the peers' parts into the
// send slots in W1's buffer), then one copy per peer.
template <typename T>
static int b0_chunk_ce(sdns_plan* p, Pipe<T>& P, const void* u_hat, bool work_layout, int c) {
    typedef typename C2<T>::type V;
    const Space& q = P.q;
    const int solver = p->cfg.solver;
    const int fam = solver == SDNS_NS ? FAM_NS_B0 : (solver == SDNS_VV ? FAM_VV_B0 : FAM_PLAIN_BWD);
    const int nfin = solver == SDNS_MHD ? 6 : 3;
    Rng mem[2], kept; k1_chunk(p, q, c, p->nchunk, &kept, mem);
    int e;
    P.st = p->stream;
    if (kept.b > kept.a)
        if ((e = P.b0(fam, reinterpret_cast<const V*>(u_hat), nfin, 0, work_layout, Rng{-1, 0}, kept, true))) return e;
    if (p->xmode == 1) CUDA_TRY(cudaEventRecord(p->ev_k[c], p->stream));
    if (kept.b <= kept.a) return SDNS_OK;
    const size_t cs = p->cs;
    const void* srcs[SDNS_XD_MAX]; void* dsts[SDNS_XD_MAX];
    for (int k = 1; k < p->P; ++k) {
        const int r = (p->rank + k) % p->P;
        const V* src = P.B + r * P.b0_slot() + (long long)kept.a * q.K2p;
        V* dst = reinterpret_cast<V*>(p->peer_ws[r] + p->off_A) + ((long long)q.c1off + kept.a) * q.K2p;
        srcs[k - 1] = src; dsts[k - 1] = dst;
        if (p->xmode == 1)
            if ((e = copy_rows(p, r, p->ev_k[c], dst, (size_t)q.K1n * q.K2p * cs, src, (size_t)q.K1l * q.K2p * cs,
                               (size_t)(kept.b - kept.a) * q.K2p * cs, (size_t)6 * q.M0p))) return e;
    }
    if (p->xmode == 2)
        return push_xfer(p, p->P - 1, srcs, dsts, (size_t)q.K1l * q.K2p * cs, (size_t)q.K1n * q.K2p * cs,
                         (size_t)(kept.b - kept.a) * q.K2p * cs, (size_t)6 * q.M0p);
    return SDNS_OK;
}

// Multi-GPU schedule of the Vortex (NS / VV) and MHD right-hand sides with the exchange on the copy engines.
// All passes run on the plan stream; B0 and F1 are cut into chunks (axis-1 modes / x0 planes), each chunk writes
// the peers' parts into send slots and, as soon as it has finished, one strided copy per peer moves them over
// NVLink on that peer's copy stream -- underneath the next chunk's passes (B0 chunks alternate with the previous
// stage's F0 chunks, F1 chunks with the Z chunks).  Before B1 and before F0 the plan stream waits for its own
// copies and meets the other ranks at the flag barrier.
template <typename T>
static int rhs_ce(sdns_plan* p, const void* u_hat, double nu, double eta, const StageOut& so) {
    typedef typename C2<T>::type V;
    Pipe<T> P(p, SDNS_SPACE_TP);
    const Space& q = P.q;
    const int solver = p->cfg.solver;
    const int nprod = solver == SDNS_MHD ? 9 : 3;
    const int nc = p->nchunk;
    const size_t cs = p->cs;
    int e;
    const bool pre = p->b0_preissued && p->b0_pre_u == u_hat && so.in_work_layout;
    p->b0_preissued = false;
    if (!pre)
        for (int c = 0; c < nc; ++c) if ((e = b0_chunk_ce<T>(p, P, u_hat, so.in_work_layout, c))) return e;
    if ((e = join_copies(p))) return e;
    P.st = p->stream;
    if ((e = P.b1(6))) return e;
    for (int c = 0; c < nc; ++c) {
        const Rng x0{chunk_bound(q.M0l, nc, c), chunk_bound(q.M0l, nc, c + 1)};
        if (x0.b <= x0.a) continue;
        if ((e = P.z(solver == SDNS_MHD ? FAM_Z_MHD : FAM_Z_CROSS, P.B, P.A, 6, true, true, x0))) return e;
        if ((e = P.f1(nprod, nullptr, x0, true))) return e;
        if (p->xmode == 1) CUDA_TRY(cudaEventRecord(p->ev_k[c], p->stream));
        const void* srcs[SDNS_XD_MAX]; void* dsts[SDNS_XD_MAX];
        for (int k = 1; k < p->P; ++k) {
            const int r = (p->rank + k) % p->P;
            const V* src = P.send_f1() + r * P.f1_slot(nprod) + (long long)x0.a * p->Nhp;
            V* dst = reinterpret_cast<V*>(p->peer_ws[r] + p->off_C) + ((long long)q.s0 + x0.a) * p->Nhp;
            srcs[k - 1] = src; dsts[k - 1] = dst;
            if (p->xmode == 1)
                if ((e = copy_rows(p, r, p->ev_k[c], dst, (size_t)q.M[0] * p->Nhp * cs, src, (size_t)q.M0l * p->Nhp * cs,
                                   (size_t)(x0.b - x0.a) * p->Nhp * cs, (size_t)nprod * p->N1p))) return e;
        }
        if (p->xmode == 2)
            if ((e = push_xfer(p, p->P - 1, srcs, dsts, (size_t)q.M0l * p->Nhp * cs, (size_t)q.M[0] * p->Nhp * cs,
                               (size_t)(x0.b - x0.a) * p->Nhp * cs, (size_t)nprod * p->N1p))) return e;
    }
    if ((e = join_copies(p))) return e;
    for (int c = 0; c < nc; ++c) {
        Rng mem[2], kept; k1_chunk(p, q, c, nc, &kept, mem);
        P.st = p->stream;
        for (int i = 0; i < 2; ++i)
            if (mem[i].b > mem[i].a)
                if ((e = launch_f0<T>(p, P, u_hat, nu, eta, so, nprod, false, (T)1, Rng{-1, 0}, mem[i]))) return e;
        if (so.next_u) if ((e = b0_chunk_ce<T>(p, P, so.next_u, true, c))) return e;
    }
    if (so.next_u) { p->b0_preissued = true; p->b0_pre_u = so.next_u; }
    const Rng tr = k1_truncated(p, q);
    P.st = p->stream;
    if (tr.b > tr.a) return launch_f0<T>(p, P, u_hat, nu, eta, so, nprod, false, (T)1, Rng{-1, 0}, tr);
    return SDNS_OK;
}

template <typename T>
static int rhs_t(sdns_plan* p, const void* u_hat, double nu, double eta, const StageOut& so) {
    typedef typename C2<T>::type V;
    Pipe<T> P(p, SDNS_SPACE_TP);
    const V* u = reinterpret_cast<const V*>(u_hat);
    int e;
    const int solver = p->cfg.solver;
    const int conv = p->cfg.convection;
    int nprod = solver == SDNS_MHD ? 9 : 3;
    bool divform = false;
    T f0scale = (T)1;
    if (solver == SDNS_NS && conv != SDNS_CONV_VORTEX) {
        V* D = reinterpret_cast<V*>(p->ws + p->off_D);
        V* S = reinterpret_cast<V*>(p->ws + p->off_S);
        if (conv == SDNS_CONV_STANDARD || conv == SDNS_CONV_SKEWED) {
            // u_j du_i/dx_j (NS.py:138-145): per component i six backward transforms (u, grad u_i) and one product
            const long long dfs = (long long)P.q.M0l * P.q.M[1] * p->Nhp;
            for (int i = 0; i < 3; ++i) {
                if ((e = P.b0(FAM_NS_GRAD_B0, u, 3, i, so.in_work_layout))) return e;
                if ((e = xbarrier(p))) return e;
                if ((e = P.b1(6))) return e;
                if ((e = xbarrier(p))) return e;           // W0 is free again on every rank before the next B0 stores into it
                if ((e = P.z(FAM_Z_DOT, P.B, D + i * dfs, 6, true, true))) return e;
            }
            if ((e = P.f1(3, D))) return e;
            if ((e = xbarrier(p))) return e;
            if (conv == SDNS_CONV_STANDARD) {
                f0scale = (T)-1;                                   // rhs = -conv (NS.py:170)
            } else {
                // Skewed (NS.py:184-189): keep the standard term in spectral space, add it in the divergence epilogue
                StridedArgs<T> a0; P.f0_geom(a0, 3);
                a0.out_mode = OUT_CONV; a0.rhs = S; a0.u_hat = u;
                if ((e = do_launch(p, p->stream, FAM_NS_F0, P.q.M[0], &a0))) return e;
            }
        }
        if (conv == SDNS_CONV_DIVERGENCE || conv == SDNS_CONV_SKEWED) {
            // d/dx_j (u_i u_j) (NS.py:147-162): three backward, six forward transforms
            if ((e = P.b0(FAM_PLAIN_BWD, u, 3, 0, so.in_work_layout))) return e;
            if ((e = xbarrier(p))) return e;
            if ((e = P.b1(3))) return e;
            if ((e = P.z(FAM_Z_UU, P.B, P.A, 3, true, true))) return e;
            if ((e = P.f1(6))) return e;
            if ((e = xbarrier(p))) return e;
            nprod = 6; divform = true;
        }
    } else {
        if (p->xmode) return rhs_ce<T>(p, u_hat, nu, eta, so);
        if (solver == SDNS_NS) { if ((e = P.b0(FAM_NS_B0, u, 3, 0, so.in_work_layout))) return e; }
        else if (solver == SDNS_VV) { if ((e = P.b0(FAM_VV_B0, u, 3, 0, so.in_work_layout))) return e; }
        else { if ((e = P.b0(FAM_PLAIN_BWD, u, 6, 0, so.in_work_layout))) return e; }
        if ((e = xbarrier(p))) return e;
        if ((e = P.b1(6))) return e;
        if ((e = P.z(solver == SDNS_MHD ? FAM_Z_MHD : FAM_Z_CROSS, P.B, P.A, 6, true, true))) return e;
        if ((e = P.f1(nprod))) return e;
        if ((e = xbarrier(p))) return e;
    }
    return launch_f0<T>(p, P, u_hat, nu, eta, so, nprod, divform, f0scale, Rng{-1, 0});
}

extern "C" int sdns_forward(sdns_plan* p, int space, int nc, const void* in, void* out) {
    int e = need_ws(p); if (e) return e;
    if (space < 0 || space > 1 || nc < 1 || !in || !out) return fail(SDNS_ERR_ARG, "sdns_forward: bad argument");
    return p->prec ? forward_t<double>(p, space, nc, in, out) : forward_t<float>(p, space, nc, in, out);
}
extern "C" int sdns_backward(sdns_plan* p, int space, int nc, const void* in, void* out) {
    int e = need_ws(p); if (e) return e;
    if (space < 0 || space > 1 || nc < 1 || !in || !out) return fail(SDNS_ERR_ARG, "sdns_backward: bad argument");
    return p->prec ? backward_t<double>(p, space, nc, in, out) : backward_t<float>(p, space, nc, in, out);
}

extern "C" int sdns_compute_rhs(sdns_plan* p, void* rhs, const void* u_hat, double nu, double eta,
                                const void* source, void* p_hat) {
    int e = need_ws(p); if (e) return e;
    if (!rhs || !u_hat) return fail(SDNS_ERR_ARG, "sdns_compute_rhs: null array");
    StageOut so; memset(&so, 0, sizeof so);
    so.out_mode = OUT_RHS; so.rhs = rhs; so.source = source; so.p_hat = p_hat;
    return p->prec ? rhs_t<double>(p, u_hat, nu, eta, so) : rhs_t<float>(p, u_hat, nu, eta, so);
}

extern "C" int sdns_compute_conv(sdns_plan* p, void* rhs, const void* u_hat) {
    int e = need_ws(p); if (e) return e;
    if (!rhs || !u_hat) return fail(SDNS_ERR_ARG, "sdns_compute_conv: null array");
    StageOut so; memset(&so, 0, sizeof so);
    so.out_mode = OUT_CONV; so.rhs = rhs;
    return p->prec ? rhs_t<double>(p, u_hat, 0, 0, so) : rhs_t<float>(p, u_hat, 0, 0, so);
}

// a, b of maths/integrators.py:185-186 in context.float, products a[rk]*dt, b[rk]*dt in that type
template <typename T>
static void rk_coeffs(int rk, double dt, double* adt, double* bdt) {
    const T a[4] = {(T)(1. / 6.), (T)(1. / 3.), (T)(1. / 3.), (T)(1. / 6.)};
    const T b[3] = {(T)0.5, (T)0.5, (T)1.};
    const T d = (T)dt;
    *adt = (double)(T)(a[rk] * d);
    *bdt = rk < 3 ? (double)(T)(b[rk] * d) : 0.0;
}

static int rk4_step_eager(sdns_plan* p, void* u_hat, void* u1, void* u2, double dt, double nu, double eta, const void* source);

// SDNS_GRAPH=1: the first step with a given argument set runs eagerly (it also initialises the kernels' attributes),
// the second one is captured on the library's stream, later ones replay the graph.
static int rk4_step_graph(sdns_plan* p, void* u_hat, void* u1, void* u2, double dt, double nu, double eta, const void* source) {
    const sdns_plan::GKey key = {u_hat, u1, u2, source, dt, nu, eta};
    if (!p->gwarm || memcmp(&key, &p->gkey, sizeof key)) {
        p->gkey = key; p->gwarm = true; p->ghave = false;
        if (p->gexec) { cudaGraphExecDestroy(p->gexec); p->gexec = nullptr; }
        return rk4_step_eager(p, u_hat, u1, u2, dt, nu, eta, source);
    }
    cudaStream_t user = p->stream;
    CUDA_TRY(cudaEventRecord(p->ev_gin, user));
    CUDA_TRY(cudaStreamWaitEvent(p->gstream, p->ev_gin, 0));
    if (!p->ghave) {
        cudaGraph_t graph = nullptr;
        const long long l0 = p->launches;
        p->stream = p->gstream; p->capturing = true;
        cudaError_t ce = cudaStreamBeginCapture(p->gstream, cudaStreamCaptureModeRelaxed);
        int e = ce == cudaSuccess ? rk4_step_eager(p, u_hat, u1, u2, dt, nu, eta, source) : SDNS_ERR_CUDA;
        if (ce == cudaSuccess) ce = cudaStreamEndCapture(p->gstream, &graph);
        p->stream = user; p->capturing = false;
        if (e) return e;
        if (ce != cudaSuccess) return fail(SDNS_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
        CUDA_TRY(cudaGraphInstantiate(&p->gexec, graph, 0));
        CUDA_TRY(cudaGraphDestroy(graph));
        p->glaunches = p->launches - l0; p->launches = l0;
        p->ghave = true;
    }
    CUDA_TRY(cudaGraphLaunch(p->gexec, p->gstream));
    p->launches += p->glaunches;
    CUDA_TRY(cudaEventRecord(p->ev_gout, p->gstream));
    CUDA_TRY(cudaStreamWaitEvent(user, p->ev_gout, 0));
#ifdef SDNS_HOST_SHIM
    p->ghave = false;                 // the emulated runtime executes a "capture" eagerly and cannot replay it
    if (p->gexec) { cudaGraphExecDestroy(p->gexec); p->gexec = nullptr; }
#endif
    return SDNS_OK;
}

extern "C" int sdns_rk4_step(sdns_plan* p, void* u_hat, void* u1, void* u2, double dt, double nu,
                             double eta, const void* source) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || !u1 || !u2) return fail(SDNS_ERR_ARG, "sdns_rk4_step: null array");
    if (p->use_graph && p->P > 1 && p->xmode >= 1 && !p->prof) return rk4_step_graph(p, u_hat, u1, u2, dt, nu, eta, source);
    return rk4_step_eager(p, u_hat, u1, u2, dt, nu, eta, source);
}

static int rk4_step_eager(sdns_plan* p, void* u_hat, void* u1, void* u2, double dt, double nu, double eta, const void* source) {
    int e;
    p->b0_preissued = false;            // nothing pre-enqueued survives from an earlier (possibly failed) call
    // Between stages the state lives in the workspace in the k1-major work layout (u1, u2 too): the
    // axis-0 passes then read and write it with a stride of one k2 row instead of N1*Nh elements.
    // Stage 0 reads the caller's u_hat (reference layout), stage 3 writes it back in that layout.
    void* u0w = p->ws + p->off_U;
    for (int rk = 0; rk < 4; ++rk) {
        StageOut so; memset(&so, 0, sizeof so);
        so.out_mode = OUT_STAGE; so.u1 = u1; so.u2 = u2; so.source = source; so.rk = rk;
        so.u0 = rk < 3 ? u0w : u_hat;
        so.in_work_layout = rk > 0;
        so.next_u = rk < 3 ? u0w : nullptr;
        if (p->prec) rk_coeffs<double>(rk, dt, &so.adt, &so.bdt); else rk_coeffs<float>(rk, dt, &so.adt, &so.bdt);
        const void* uin = rk == 0 ? u_hat : u0w;
        e = p->prec ? rhs_t<double>(p, uin, nu, eta, so) : rhs_t<float>(p, uin, nu, eta, so);
        if (e) return e;
    }
    return SDNS_OK;
}

// ---- small elementwise kernels (integrators other than RK4, cross2) ---------------------
template <typename T>
__global__ void euler_kernel(typename C2<T>::type* u, const typename C2<T>::type* r, T dt, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        typename C2<T>::type a = u[i], b = r[i];
        a.x += b.x * dt; a.y += b.y * dt; u[i] = a;
    }
}
// AB2 (integrators.py:167-175): u0 += rhs*dt (tstep 0) or 1.5*rhs*dt - 0.5*u1 ; u1 = rhs*dt
template <typename T>
__global__ void ab2_kernel(typename C2<T>::type* u, typename C2<T>::type* u1, const typename C2<T>::type* r,
                           T dt, int first, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        typename C2<T>::type a = u[i], b = r[i], o = u1[i];
        const T rx = b.x * dt, ry = b.y * dt;
        if (first) { a.x += rx; a.y += ry; }
        else { a.x += ((T)1.5 * rx - (T)0.5 * o.x); a.y += ((T)1.5 * ry - (T)0.5 * o.y); }
        u[i] = a; o.x = rx; o.y = ry; u1[i] = o;
    }
}
template <typename T>
__global__ void cross2_kernel(typename C2<T>::type* c, const typename C2<T>::type* b, const T* kx, const T* ky,
                              const T* kz, int N0, int N1, int Nh, int over_k2) {
    typedef typename C2<T>::type V;
    const long long n = (long long)N0 * N1 * Nh;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % Nh); const long long r = i / Nh;
        const int i1 = (int)(r % N1); const int i0 = (int)(r / N1);
        T k0 = kx[i0], k1 = ky[i1], k2 = kz[i2];
        if (over_k2) {
            T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;
            if (ksq == (T)0) ksq = (T)1;
            k0 /= ksq; k1 /= ksq; k2 /= ksq;
        }
        const V b0 = b[i], b1 = b[n + i], b2 = b[2 * n + i];
        c[i] = icross<T, V>(k1, b2, k2, b1);
        c[n + i] = icross<T, V>(k2, b0, k0, b2);
        c[2 * n + i] = icross<T, V>(k0, b1, k1, b0);
    }
}

static int ncomp_state(const sdns_plan* p) { return p->cfg.solver == SDNS_MHD ? 6 : 3; }

extern "C" int sdns_euler_step(sdns_plan* p, void* u_hat, void* rhs, double dt, double nu, double eta,
                               const void* source) {
    int e = sdns_compute_rhs(p, rhs, u_hat, nu, eta, source, nullptr); if (e) return e;
    const long long n = (long long)ncomp_state(p) * p->N[0] * p->N1l * p->Nh;
    if (p->prec) SDNS_LAUNCH(euler_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)u_hat, (const double2*)rhs, dt, n);
    else SDNS_LAUNCH(euler_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)u_hat, (const float2*)rhs, (float)dt, n);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}

extern "C" int sdns_ab2_step(sdns_plan* p, void* u_hat, void* u1, void* rhs, double dt, int tstep,
                             double nu, double eta, const void* source) {
    int e = sdns_compute_rhs(p, rhs, u_hat, nu, eta, source, nullptr); if (e) return e;
    const long long n = (long long)ncomp_state(p) * p->N[0] * p->N1l * p->Nh;
    if (p->prec) SDNS_LAUNCH(ab2_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)u_hat, (double2*)u1, (const double2*)rhs, dt, tstep == 0, n);
    else SDNS_LAUNCH(ab2_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)u_hat, (float2*)u1, (const float2*)rhs, (float)dt, tstep == 0, n);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}

extern "C" int sdns_cross2(sdns_plan* p, void* c, const void* b, int over_k2) {
    int e = need_ws(p); if (e) return e;
    if (!c || !b || c == b) return fail(SDNS_ERR_ARG, "sdns_cross2: c and b must be distinct arrays");
    if (p->prec)
        SDNS_LAUNCH(cross2_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)c, (const double2*)b,
            (const double*)(p->ws + p->kx_off), (const double*)(p->ws + p->ky_off), (const double*)(p->ws + p->kz_off),
            p->N[0], p->N1l, p->Nh, over_k2);
    else
        SDNS_LAUNCH(cross2_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)c, (const float2*)b,
            (const float*)(p->ws + p->kx_off), (const float*)(p->ws + p->ky_off), (const float*)(p->ws + p->kz_off),
            p->N[0], p->N1l, p->Nh, over_k2);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}

extern "C" int sdns_energy(sdns_plan* p, const void* u_hat, int nc, double* out) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || !out || nc < 1) return fail(SDNS_ERR_ARG, "sdns_energy: bad argument");
    const long long n = (long long)nc * p->N[0] * p->N1l * p->Nh;
    double* red = reinterpret_cast<double*>(p->ws + p->off_red);
    const int nb = p->red_blocks;
    if (p->prec) SDNS_LAUNCH(energy_kernel<double>, nb, 256, 0, p->stream)((const double2*)u_hat, n, p->Nh, p->N[2], red);
    else SDNS_LAUNCH(energy_kernel<float>, nb, 256, 0, p->stream)((const float2*)u_hat, n, p->Nh, p->N[2], red);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    std::vector<double> h(nb);
    CUDA_TRY(cudaMemcpyAsync(h.data(), red, sizeof(double) * nb, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
    *out = s;
    return check_comm(p);
}

extern "C" int sdns_rk4_steps_host(sdns_plan* p, void* host_u, void* du, void* d1, void* d2, int nsteps,
                                   double dt, double nu, double eta) {
    int e = need_ws(p); if (e) return e;
    if (!host_u || !du || !d1 || !d2 || nsteps < 0) return fail(SDNS_ERR_ARG, "sdns_rk4_steps_host: bad argument");
    const size_t bytes = (size_t)ncomp_state(p) * p->N[0] * p->N1l * p->Nh * p->cs;
    CUDA_TRY(cudaMemcpyAsync(du, host_u, bytes, cudaMemcpyHostToDevice, p->stream));
    for (int s = 0; s < nsteps; ++s) { e = sdns_rk4_step(p, du, d1, d2, dt, nu, eta, nullptr); if (e) return e; }
    CUDA_TRY(cudaMemcpyAsync(host_u, du, bytes, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return check_comm(p);
}

// ---- diagnostics and forcing of demo/Isotropic.py on the device (update(): :161-187, :220-259; spectrum(): :88-118).
// The reference computes them with numpy on the host arrays of the context; here the state stays on the GPU and each
// quantity is one reduction kernel over the local spectral block (the caller reduces over ranks).
enum DiagMode { DIAG_ENERGY_W = 0, DIAG_ENSTROPHY = 1, DIAG_DIVERGENCE = 2 };
// DIAG_ENERGY_W : sum_c w_h |u_c * weight|^2               energy_fourier(U_hat*k2_mask, T)      (Isotropic.py:168)
// DIAG_ENSTROPHY: sum_c w_h |(i K x u)_c|^2                 energy_fourier(cross2(K, U_hat), T)   (Isotropic.py:243-244)
// DIAG_DIVERGENCE: w_h |i K.u|^2                            L2_norm(get_divergence) by Parseval   (Isotropic.py:245-247)
// w_h = 1 on the k2 = 0 and k2 = N2/2 planes, 2 elsewhere (Hermitian half spectrum)
template <typename T, typename W, int MODE>
__global__ void diag_kernel(const typename C2<T>::type* __restrict__ u, const W* __restrict__ weight, long long n1, int ncomp,
                            const T* kx, const T* ky, const T* kz, int N1, int Nh, int N2, double* __restrict__ out) {
    typedef typename C2<T>::type V;
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (long long)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % Nh);
        const double wh = (i2 == 0 || 2 * i2 == N2) ? 1.0 : 2.0;
        if (MODE == DIAG_ENERGY_W) {
            const double m = weight ? (double)weight[i] : 1.0;
            double e = 0;
            for (int c = 0; c < ncomp; ++c) {
                const V v = u[c * n1 + i];
                const double px = (double)v.x * m, py = (double)v.y * m;
                e += px * px + py * py;
            }
            s += wh * e;
        } else {
            const long long r = i / Nh;
            const int i1 = (int)(r % N1), i0 = (int)(r / N1);
            const T k0 = kx[i0], k1 = ky[i1], k2 = kz[i2];
            const V a = u[i], b = u[n1 + i], c = u[2 * n1 + i];
            if (MODE == DIAG_ENSTROPHY) {
                const V w0 = icross<T, V>(k1, c, k2, b), w1 = icross<T, V>(k2, a, k0, c), w2 = icross<T, V>(k0, b, k1, a);
                s += wh * ((double)w0.x * w0.x + (double)w0.y * w0.y + (double)w1.x * w1.x + (double)w1.y * w1.y +
                           (double)w2.x * w2.x + (double)w2.y * w2.y);
            } else {
                const T dx = k0 * a.x + k1 * b.x + k2 * c.x, dy = k0 * a.y + k1 * b.y + k2 * c.y;
                s += wh * ((double)dx * dx + (double)dy * dy);
            }
        }
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

// U_hat *= a*factor + b*(1 - factor), factor a real field broadcast over the components: with (a, b) = (1, 0) the plain
// product, with (alpha, 1) and factor = k2_mask the forcing rescale of Isotropic.py:180 without forming its factor array.
// numpy forms the product in the promoted type and rounds once; so does this.
template <typename T, typename W>
__global__ void scale_field_kernel(typename C2<T>::type* u, const W* __restrict__ f, long long n1, int ncomp, double a, double b) {
    typedef typename C2<T>::type V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (long long)gridDim.x * blockDim.x) {
        const double fi = (double)f[i];
        const double m = a * fi + b * (1.0 - fi);          // (a, b) = (1, 0): the field itself; (alpha, 1): alpha*mask + (1 - mask)
        for (int c = 0; c < ncomp; ++c) {
            V v = u[c * n1 + i];
            v.x = (T)((double)v.x * m); v.y = (T)((double)v.y * m);
            u[c * n1 + i] = v;
        }
    }
}
template <typename T>
__global__ void set_mode_kernel(typename C2<T>::type* u, long long n1, int ncomp, long long idx, T re, T im) {
    const int c = threadIdx.x;
    if (c < ncomp) { u[c * n1 + idx].x = re; u[c * n1 + idx].y = im; }
}

#ifdef SDNS_HOST_SHIM
static inline void sdns_atomic_add(double* p, double v) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = __atomic_load_n(q, __ATOMIC_RELAXED), nw;
    do { double d; memcpy(&d, &old, 8); d += v; memcpy(&nw, &d, 8); } while (!__atomic_compare_exchange_n(q, &old, nw, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
#else
static __device__ __forceinline__ void sdns_atomic_add(double* p, double v) { atomicAdd(p, v); }
#endif

// Shell sums of spectrum() (Isotropic.py:88-118): uiui = w_h * sum_c |u_c|^2 (the 4 pi / 3 factor, the shell volumes
// k^3 - k0^3 and the division by the point counts are the caller's, after the reduction over ranks); shell i collects
// the modes with i + 0.5 < |k| <= i + 1.5 -- np.digitize(sqrt(K2), bins, right=True) == i + 1 for bins = 0.5, 1.5, ...
// hist[0 .. nbins) += uiui, hist[nbins .. 2 nbins) += 1.  Block histogram in shared memory, one flush per block.
template <typename T>
__global__ void spectrum_kernel(const typename C2<T>::type* __restrict__ u, long long n1, int ncomp, const T* kx, const T* ky,
                                const T* kz, int N1, int Nh, int N2, int nbins, double* __restrict__ hist) {
    typedef typename C2<T>::type V;
    SDNS_DYN_SMEM(smraw);
    double* sh = reinterpret_cast<double*>(smraw);
    for (int b = threadIdx.x; b < 2 * nbins; b += blockDim.x) sh[b] = 0.0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (long long)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % Nh);
        const long long r = i / Nh;
        const int i1 = (int)(r % N1), i0 = (int)(r / N1);
        const T k0 = kx[i0], k1 = ky[i1], k2 = kz[i2];
        T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;            // K2 as get_context builds it (NS.py:42-44)
        const double km = sqrt((double)ksq);
        // digitize(km, bins, right=True): smallest z with km <= z + 0.5
        const int z = km <= 0.5 ? 0 : (int)ceil(km - 0.5);
        const int shell = z - 1;
        if (shell < 0 || shell >= nbins - 1) continue;
        double e = 0;
        for (int c = 0; c < ncomp; ++c) { const V v = u[c * n1 + i]; e += (double)v.x * v.x + (double)v.y * v.y; }
        const double wh = (i2 == 0 || i2 == Nh - 1) ? 1.0 : 2.0;      // Isotropic.py:90-93: planes 0 and -1 once, the others twice
        sdns_atomic_add(&sh[shell], wh * e);
        sdns_atomic_add(&sh[nbins + shell], 1.0);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 2 * nbins; b += blockDim.x) if (sh[b] != 0.0) sdns_atomic_add(&hist[b], sh[b]);
}

static int reduce_blocks(sdns_plan* p, double* out) {
    const int nb = p->red_blocks;
    std::vector<double> h(nb);
    CUDA_TRY(cudaMemcpyAsync(h.data(), p->ws + p->off_red, sizeof(double) * nb, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
    *out = s;
    return check_comm(p);
}

template <typename T, typename W, int MODE>
static void launch_diag(sdns_plan* p, const void* u, const void* w, int ncomp) {
    typedef typename C2<T>::type V;
    const long long n1 = (long long)p->N[0] * p->N1l * p->Nh;
    SDNS_LAUNCH((diag_kernel<T, W, MODE>), p->red_blocks, 256, 0, p->stream)((const V*)u, (const W*)w, n1, ncomp,
        (const T*)(p->ws + p->kx_off), (const T*)(p->ws + p->ky_off), (const T*)(p->ws + p->kz_off), p->N1l, p->Nh, p->N[2],
        reinterpret_cast<double*>(p->ws + p->off_red));
    p->launches++;
}

extern "C" int sdns_energy_weighted(sdns_plan* p, const void* u_hat, int ncomp, const void* weight, int weight_is_double, double* out) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || !out || ncomp < 1) return fail(SDNS_ERR_ARG, "sdns_energy_weighted: bad argument");
    if (p->prec) { if (weight_is_double) launch_diag<double, double, DIAG_ENERGY_W>(p, u_hat, weight, ncomp); else launch_diag<double, float, DIAG_ENERGY_W>(p, u_hat, weight, ncomp); }
    else { if (weight_is_double) launch_diag<float, double, DIAG_ENERGY_W>(p, u_hat, weight, ncomp); else launch_diag<float, float, DIAG_ENERGY_W>(p, u_hat, weight, ncomp); }
    CUDA_TRY(cudaGetLastError());
    return reduce_blocks(p, out);
}
extern "C" int sdns_enstrophy(sdns_plan* p, const void* u_hat, double* out) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || !out) return fail(SDNS_ERR_ARG, "sdns_enstrophy: bad argument");
    if (p->prec) launch_diag<double, double, DIAG_ENSTROPHY>(p, u_hat, nullptr, 3); else launch_diag<float, float, DIAG_ENSTROPHY>(p, u_hat, nullptr, 3);
    CUDA_TRY(cudaGetLastError());
    return reduce_blocks(p, out);
}
extern "C" int sdns_divergence_norm(sdns_plan* p, const void* u_hat, double* out) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || !out) return fail(SDNS_ERR_ARG, "sdns_divergence_norm: bad argument");
    if (p->prec) launch_diag<double, double, DIAG_DIVERGENCE>(p, u_hat, nullptr, 3); else launch_diag<float, float, DIAG_DIVERGENCE>(p, u_hat, nullptr, 3);
    CUDA_TRY(cudaGetLastError());
    return reduce_blocks(p, out);
}
extern "C" int sdns_scale_field(sdns_plan* p, void* u_hat, int ncomp, const void* factor, int factor_is_double, double a, double b) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || !factor || ncomp < 1) return fail(SDNS_ERR_ARG, "sdns_scale_field: bad argument");
    const long long n1 = (long long)p->N[0] * p->N1l * p->Nh;
    if (p->prec) {
        if (factor_is_double) SDNS_LAUNCH((scale_field_kernel<double, double>), SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)u_hat, (const double*)factor, n1, ncomp, a, b);
        else SDNS_LAUNCH((scale_field_kernel<double, float>), SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)u_hat, (const float*)factor, n1, ncomp, a, b);
    } else {
        if (factor_is_double) SDNS_LAUNCH((scale_field_kernel<float, double>), SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)u_hat, (const double*)factor, n1, ncomp, a, b);
        else SDNS_LAUNCH((scale_field_kernel<float, float>), SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)u_hat, (const float*)factor, n1, ncomp, a, b);
    }
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}
extern "C" int sdns_set_mode(sdns_plan* p, void* u_hat, int ncomp, int i0, int i1, int i2, double re, double im) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || ncomp < 1 || ncomp > 32 || i0 < 0 || i0 >= p->N[0] || i1 < 0 || i1 >= p->N1l || i2 < 0 || i2 >= p->Nh)
        return fail(SDNS_ERR_ARG, "sdns_set_mode: bad argument");
    const long long n1 = (long long)p->N[0] * p->N1l * p->Nh;
    const long long idx = ((long long)i0 * p->N1l + i1) * p->Nh + i2;
    if (p->prec) SDNS_LAUNCH(set_mode_kernel<double>, 1, 32, 0, p->stream)((double2*)u_hat, n1, ncomp, idx, re, im);
    else SDNS_LAUNCH(set_mode_kernel<float>, 1, 32, 0, p->stream)((float2*)u_hat, n1, ncomp, idx, (float)re, (float)im);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}
extern "C" int sdns_spectrum(sdns_plan* p, const void* u_hat, int ncomp, int nbins, double* sums, double* counts) {
    int e = need_ws(p); if (e) return e;
    if (!u_hat || !sums || !counts || ncomp < 1 || nbins < 2 || nbins > SDNS_MAX_BINS) return fail(SDNS_ERR_ARG, "sdns_spectrum: bad argument (2 <= nbins <= 4096)");
    double* hist = reinterpret_cast<double*>(p->ws + p->off_red);
    CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * 2 * nbins, p->stream));
    const long long n1 = (long long)p->N[0] * p->N1l * p->Nh;
    const size_t smem = sizeof(double) * 2 * nbins;
#ifndef SDNS_HOST_SHIM
    static bool once = false;
    if (!once) {
        CUDA_TRY(cudaFuncSetAttribute(spectrum_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 2 * SDNS_MAX_BINS)));
        CUDA_TRY(cudaFuncSetAttribute(spectrum_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 2 * SDNS_MAX_BINS)));
        once = true;
    }
#endif
    if (p->prec) SDNS_LAUNCH(spectrum_kernel<double>, p->red_blocks, 256, smem, p->stream)((const double2*)u_hat, n1, ncomp,
        (const double*)(p->ws + p->kx_off), (const double*)(p->ws + p->ky_off), (const double*)(p->ws + p->kz_off), p->N1l, p->Nh, p->N[2], nbins, hist);
    else SDNS_LAUNCH(spectrum_kernel<float>, p->red_blocks, 256, smem, p->stream)((const float2*)u_hat, n1, ncomp,
        (const float*)(p->ws + p->kx_off), (const float*)(p->ws + p->ky_off), (const float*)(p->ws + p->kz_off), p->N1l, p->Nh, p->N[2], nbins, hist);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    std::vector<double> h(2 * nbins);
    CUDA_TRY(cudaMemcpyAsync(h.data(), hist, sizeof(double) * 2 * nbins, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (int i = 0; i < nbins; ++i) { sums[i] = h[i]; counts[i] = h[nbins + i]; }
    return check_comm(p);
}

// ---- standalone pointwise operators of the fine-grained plug-in surface (optimization/__init__.py:12-55)
template <typename T>
__global__ void cross1_kernel(T* c, const T* a, const T* b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T a0 = a[i], a1 = a[n + i], a2 = a[2 * n + i];
        const T b0 = b[i], b1 = b[n + i], b2 = b[2 * n + i];
        c[i] = a1 * b2 - a2 * b1; c[n + i] = a2 * b0 - a0 * b2; c[2 * n + i] = a0 * b1 - a1 * b0;
    }
}
template <typename T>
__global__ void cross2_dense_kernel(typename C2<T>::type* c, const T* a, const typename C2<T>::type* b, long long n) {
    typedef typename C2<T>::type V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T a0 = a[i], a1 = a[n + i], a2 = a[2 * n + i];
        const V b0 = b[i], b1 = b[n + i], b2 = b[2 * n + i];
        c[i] = icross<T, V>(a1, b2, a2, b1);
        c[n + i] = icross<T, V>(a2, b0, a0, b2);
        c[2 * n + i] = icross<T, V>(a0, b1, a1, b0);
    }
}
// project(u, K, K_over_K2): u -= sum(K_over_K2*u, 0)*K   (maths/maths.py:8-11)
template <typename T>
__global__ void project_kernel(typename C2<T>::type* u, const T* kx, const T* ky, const T* kz, int N0, int N1, int Nh) {
    typedef typename C2<T>::type V;
    const long long n = (long long)N0 * N1 * Nh;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % Nh); const long long r = i / Nh;
        const int i1 = (int)(r % N1); const int i0 = (int)(r / N1);
        const T k0 = kx[i0], k1 = ky[i1], k2 = kz[i2];
        T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;
        const T ks = ksq == (T)0 ? (T)1 : ksq;
        V u0 = u[i], u1 = u[n + i], u2 = u[2 * n + i];
        V p;
        p.x = u0.x * (k0 / ks) + u1.x * (k1 / ks); p.x += u2.x * (k2 / ks);
        p.y = u0.y * (k0 / ks) + u1.y * (k1 / ks); p.y += u2.y * (k2 / ks);
        u0.x -= p.x * k0; u0.y -= p.y * k0; u1.x -= p.x * k1; u1.y -= p.y * k1; u2.x -= p.x * k2; u2.y -= p.y * k2;
        u[i] = u0; u[n + i] = u1; u[2 * n + i] = u2;
    }
}

// add_pressure_diffusion_NS(du, u_hat, nu, ksq, kk, p_hat, k_over_k2)  (solvers/NS.py:203-217, cython_solvers.in:40-80):
// p_hat = sum_i du_i K_i/K^2 ; du_i -= p_hat K_i + nu K^2 u_hat_i -- the same arithmetic, in the same order, as the F0
// epilogue that fuses it into the right-hand side
template <typename T>
__global__ void pressure_diffusion_kernel(typename C2<T>::type* du, const typename C2<T>::type* uh, typename C2<T>::type* ph,
                                          T nu, const T* kx, const T* ky, const T* kz, int N0, int N1, int Nh) {
    typedef typename C2<T>::type V;
    const long long n = (long long)N0 * N1 * Nh;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int i2 = (int)(i % Nh); const long long r = i / Nh;
        const int i1 = (int)(r % N1); const int i0 = (int)(r / N1);
        const T k0 = kx[i0], k1 = ky[i1], k2 = kz[i2];
        T ksq = k0 * k0; ksq += k1 * k1; ksq += k2 * k2;
        const T ks = ksq == (T)0 ? (T)1 : ksq;
        const T q0 = k0 / ks, q1 = k1 / ks, q2 = k2 / ks;
        const T z = nu * ksq;
        V d0 = du[i], d1 = du[n + i], d2 = du[2 * n + i];
        const V w0 = uh[i], w1 = uh[n + i], w2 = uh[2 * n + i];
        V p;
        p.x = d0.x * q0 + d1.x * q1; p.x += d2.x * q2;
        p.y = d0.y * q0 + d1.y * q1; p.y += d2.y * q2;
        if (ph) ph[i] = p;
        d0.x -= p.x * k0; d0.y -= p.y * k0; d1.x -= p.x * k1; d1.y -= p.y * k1; d2.x -= p.x * k2; d2.y -= p.y * k2;
        d0.x -= z * w0.x; d0.y -= z * w0.y; d1.x -= z * w1.x; d1.y -= z * w1.y; d2.x -= z * w2.x; d2.y -= z * w2.y;
        du[i] = d0; du[n + i] = d1; du[2 * n + i] = d2;
    }
}

extern "C" int sdns_add_pressure_diffusion(sdns_plan* p, void* du, const void* u_hat, double nu, void* p_hat) {
    int e = need_ws(p); if (e) return e;
    if (!du || !u_hat || du == u_hat) return fail(SDNS_ERR_ARG, "sdns_add_pressure_diffusion: du and u_hat must be distinct arrays");
    if (p->prec) SDNS_LAUNCH(pressure_diffusion_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)du, (const double2*)u_hat, (double2*)p_hat, nu,
        (const double*)(p->ws + p->kx_off), (const double*)(p->ws + p->ky_off), (const double*)(p->ws + p->kz_off), p->N[0], p->N1l, p->Nh);
    else SDNS_LAUNCH(pressure_diffusion_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)du, (const float2*)u_hat, (float2*)p_hat, (float)nu,
        (const float*)(p->ws + p->kx_off), (const float*)(p->ws + p->ky_off), (const float*)(p->ws + p->kz_off), p->N[0], p->N1l, p->Nh);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}

extern "C" int sdns_cross1(sdns_plan* p, void* c, const void* a, const void* b, long long n) {
    if (!p || !c || !a || !b || n < 1) return fail(SDNS_ERR_ARG, "sdns_cross1: bad argument");
    if (p->prec) SDNS_LAUNCH(cross1_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double*)c, (const double*)a, (const double*)b, n);
    else SDNS_LAUNCH(cross1_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float*)c, (const float*)a, (const float*)b, n);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}
extern "C" int sdns_cross2_dense(sdns_plan* p, void* c, const void* a, const void* b) {
    if (!p || !c || !a || !b || c == b) return fail(SDNS_ERR_ARG, "sdns_cross2_dense: bad argument");
    const long long n = (long long)p->N[0] * p->N1l * p->Nh;
    if (p->prec) SDNS_LAUNCH(cross2_dense_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)c, (const double*)a, (const double2*)b, n);
    else SDNS_LAUNCH(cross2_dense_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)c, (const float*)a, (const float2*)b, n);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}
extern "C" int sdns_project(sdns_plan* p, void* u) {
    int e = need_ws(p); if (e) return e;
    if (!u) return fail(SDNS_ERR_ARG, "sdns_project: null array");
    if (p->prec) SDNS_LAUNCH(project_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)u,
        (const double*)(p->ws + p->kx_off), (const double*)(p->ws + p->ky_off), (const double*)(p->ws + p->kz_off), p->N[0], p->N1l, p->Nh);
    else SDNS_LAUNCH(project_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)u,
        (const float*)(p->ws + p->kx_off), (const float*)(p->ws + p->ky_off), (const float*)(p->ws + p->kz_off), p->N[0], p->N1l, p->Nh);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}

// ---- building blocks of the embedded Runge-Kutta integrator (maths/integrators.py:15-147) ---------
struct LinArgs { const void* x[9]; double c[9]; int n; };

// out = (base ? base : 0) + sum_t c_t * x_t     (complex arrays of n elements)
template <typename T>
__global__ void lincomb_kernel(typename C2<T>::type* out, const typename C2<T>::type* base, LinArgs la, long long n) {
    typedef typename C2<T>::type V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        V s; s.x = 0; s.y = 0;
        if (base) s = base[i];
        for (int t = 0; t < la.n; ++t) {
            const V v = reinterpret_cast<const V*>(la.x[t])[i];
            const T c = (T)la.c[t];
            s.x += c * v.x; s.y += c * v.y;
        }
        out[i] = s;
    }
}

// per component: sum |err / (atol + max(|u0|,|u1|)*rtol)|^2   (integrators.py:86-92)
template <typename T>
__global__ void errnorm_kernel(const typename C2<T>::type* u0, const typename C2<T>::type* u1,
                               const typename C2<T>::type* err, T atol, T rtol, long long n, double* out) {
    typedef typename C2<T>::type V;
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const V a = u0[i], b = u1[i], e = err[i];
        const T ma = sqrt(a.x * a.x + a.y * a.y), mb = sqrt(b.x * b.x + b.y * b.y);
        const T sc = atol + (ma > mb ? ma : mb) * rtol;
        const T r = sqrt(e.x * e.x + e.y * e.y) / sc;
        s += (double)r * (double)r;
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

extern "C" int sdns_lincomb(sdns_plan* p, void* out, const void* base, int nterms, const double* coeffs,
                            const void* const* arrays, int ncomp) {
    if (!p || !out || nterms < 0 || nterms > 9 || (nterms && (!coeffs || !arrays)) || ncomp < 1)
        return fail(SDNS_ERR_ARG, "sdns_lincomb: bad argument");
    LinArgs la; la.n = nterms;
    for (int t = 0; t < nterms; ++t) { la.x[t] = arrays[t]; la.c[t] = coeffs[t]; }
    const long long n = (long long)ncomp * p->N[0] * p->N1l * p->Nh;
    if (p->prec) SDNS_LAUNCH(lincomb_kernel<double>, SDNS_EW_BLOCKS, 256, 0, p->stream)((double2*)out, (const double2*)base, la, n);
    else SDNS_LAUNCH(lincomb_kernel<float>, SDNS_EW_BLOCKS, 256, 0, p->stream)((float2*)out, (const float2*)base, la, n);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return SDNS_OK;
}

extern "C" int sdns_errnorm(sdns_plan* p, const void* u0, const void* u1, const void* err, double atol, double rtol,
                            int ncomp, double* out) {
    int e = need_ws(p); if (e) return e;
    if (!u0 || !u1 || !err || !out || ncomp < 1) return fail(SDNS_ERR_ARG, "sdns_errnorm: bad argument");
    const long long n = (long long)p->N[0] * p->N1l * p->Nh;
    double* red = reinterpret_cast<double*>(p->ws + p->off_red);
    const int nb = p->red_blocks;
    std::vector<double> h(nb);
    for (int k = 0; k < ncomp; ++k) {
        if (p->prec) SDNS_LAUNCH(errnorm_kernel<double>, nb, 256, 0, p->stream)((const double2*)u0 + k * n, (const double2*)u1 + k * n,
                                                                       (const double2*)err + k * n, atol, rtol, n, red);
        else SDNS_LAUNCH(errnorm_kernel<float>, nb, 256, 0, p->stream)((const float2*)u0 + k * n, (const float2*)u1 + k * n,
                                                              (const float2*)err + k * n, (float)atol, (float)rtol, n, red);
        p->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(h.data(), red, sizeof(double) * nb, cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        double s = 0; for (int i = 0; i < nb; ++i) s += h[i];
        out[k] = s;
    }
    return check_comm(p);
}

// ---- profiling ----------------------------------------------------------------------------
extern "C" int sdns_profile_enable(sdns_plan* p, int on) {
    if (!p) return fail(SDNS_ERR_ARG, "null plan");
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (cudaStream_t y : p->ys) if (y) CUDA_TRY(cudaStreamSynchronize(y));
    p->prof = on != 0; p->recs.clear(); p->crecs.clear(); p->brecs.clear(); p->ev_used = 0;
    p->tl_on = on == 2; p->timeline.clear();
    if (p->tl_on) {
        if (!p->tl_base) CUDA_TRY(cudaEventCreate(&p->tl_base));
        CUDA_TRY(cudaEventRecord(p->tl_base, p->stream));
    }
    for (int i = 0; i < FAM_COUNT; ++i) { p->prof_ms[i] = 0; p->prof_bytes[i] = 0; p->prof_remote[i] = 0; p->prof_n[i] = 0; }
    p->copy_bytes = 0; p->copy_n = 0; for (int i = 0; i < 32; ++i) p->copy_ms[i] = 0;
    return SDNS_OK;
}
extern "C" int sdns_profile_read(sdns_plan* p, int family, double* total_ms, long long* launches, double* bytes) {
    if (!p || family < 0 || family >= FAM_COUNT) return fail(SDNS_ERR_ARG, "sdns_profile_read: bad argument");
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (const sdns_plan::Rec& r : p->recs) {
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        p->prof_ms[r.fam] += ms; p->prof_bytes[r.fam] += r.bytes; p->prof_remote[r.fam] += r.remote; p->prof_n[r.fam]++;
    }
    for (cudaStream_t y : p->ys) if (y) CUDA_TRY(cudaStreamSynchronize(y));
    for (const sdns_plan::CRec& r : p->crecs) {
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
        p->copy_ms[r.s] += ms; p->copy_bytes += r.bytes; p->copy_n++;
    }
    if (p->tl_on) {
        auto row = [&](double kind, cudaEvent_t a, cudaEvent_t b, double bytes) {
            float t0 = 0, t1 = 0;
            if (cudaEventElapsedTime(&t0, p->tl_base, a) != cudaSuccess || cudaEventElapsedTime(&t1, p->tl_base, b) != cudaSuccess) return;
            p->timeline.push_back(kind); p->timeline.push_back(t0); p->timeline.push_back(t1); p->timeline.push_back(bytes);
        };
        for (const sdns_plan::Rec& r : p->recs) row(r.fam, r.a, r.b, r.bytes);
        for (const sdns_plan::Rec& r : p->brecs) row(r.fam == -2 ? 98 : 99, r.a, r.b, 0);
        for (const sdns_plan::CRec& r : p->crecs) row(100 + r.s, r.a, r.b, r.bytes);
    }
    p->recs.clear(); p->crecs.clear(); p->brecs.clear(); p->ev_used = 0;
    if (total_ms) *total_ms = p->prof_ms[family];
    if (launches) *launches = p->prof_n[family];
    if (bytes) *bytes = p->prof_bytes[family];
    return SDNS_OK;
}

extern "C" int sdns_profile_read_nvlink(sdns_plan* p, int family, double* bytes) {
    if (!p || !bytes || family < 0 || family >= FAM_COUNT) return fail(SDNS_ERR_ARG, "sdns_profile_read_nvlink: bad argument");
    int e = sdns_profile_read(p, family, nullptr, nullptr, nullptr); if (e) return e;
    *bytes = p->prof_remote[family];
    return SDNS_OK;
}

extern "C" int sdns_profile_read_copies(sdns_plan* p, double* busy_ms, double* bytes, long long* ncopies) {
    if (!p) return fail(SDNS_ERR_ARG, "sdns_profile_read_copies: null plan");
    int e = sdns_profile_read(p, 0, nullptr, nullptr, nullptr); if (e) return e;
    double mx = 0; for (int i = 0; i < 32; ++i) mx = std::max(mx, p->copy_ms[i]);
    if (busy_ms) *busy_ms = mx;
    if (bytes) *bytes = p->copy_bytes;
    if (ncopies) *ncopies = p->copy_n;
    return SDNS_OK;
}

extern "C" int sdns_xfer_stats(sdns_plan* p, double* bytes, long long* flushes) {
    if (!p) return fail(SDNS_ERR_ARG, "sdns_xfer_stats: null plan");
    if (bytes) *bytes = p->xfer_bytes;
    if (flushes) *flushes = p->xfer_flushes;
    return SDNS_OK;
}

extern "C" int sdns_profile_timeline(sdns_plan* p, double* rows, int max_rows, int* nrows) {
    if (!p || !nrows) return fail(SDNS_ERR_ARG, "sdns_profile_timeline: bad argument");
    int e = sdns_profile_read(p, 0, nullptr, nullptr, nullptr); if (e) return e;
    const int n = (int)(p->timeline.size() / 4);
    *nrows = n;
    if (rows) for (int i = 0; i < 4 * std::min(n, max_rows); ++i) rows[i] = p->timeline[i];
    return SDNS_OK;
}
