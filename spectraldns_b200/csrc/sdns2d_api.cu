// sdns2d_api.cu -- doubly periodic (2-D) solvers NS2D and Bq2D on the B200 path: plan object and C ABI (sdns2d_*,
// include/sdns_b200.h).
//
// Reference: solvers/NS2D.py:13-51 (vorticity form u x curl with a scalar curl), solvers/Bq2D.py:101-176 (Boussinesq:
// velocity + density, buoyancy in the pressure and in the u1 equation), their compiled forms
// optimization/cython_solvers.in:82-127 (add_pressure_diffusion_Bq2D / _NS2D) and cython_maths.in:89-147 (cross1_2D,
// cross2_2D), and the 2-D transforms shenfun runs underneath T.forward / T.backward (axis 0 c2c, axis 1 r2c).
//
// One right-hand side = five launches built from the 3-D path's kernels:
//   prep      the fields that enter the backward transform, (u0, u1[, rho], curl_hat = i (K0 u1 - K1 u0))
//   B0        axis-0 inverse c2c (strided_kernel S_PLAIN) with the 2/3 truncation or the 3/2 zero padding at load
//   Z         axis-1 c2r of all fields, the real-space products, r2c of the products (z_kernel Z_NS2D / Z_BQ2D)
//   F0        axis-0 forward c2c (+ 3/2 truncation)
//   epilogue  Nyquist mask, pressure, diffusion, buoyancy, and either the right-hand side or the RK4 stage update
// A doubly periodic problem is a thousandth of a triply periodic one; single GPU only.
#ifndef SDNS_HOST_SHIM
#include <cuda_runtime.h>
#endif
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/sdns_b200.h"
#include "launch.cuh"

using namespace sdns;

typedef int (*launch_fn)(int, const void*, cudaStream_t);
extern launch_fn g_launch[FAM_COUNT][2];

#ifdef SDNS_HOST_SHIM
#define SDNS2D_EW_BLOCKS 2
#else
#define SDNS2D_EW_BLOCKS 592
#endif

static thread_local std::string g_err2;
extern "C" const char* sdns2d_last_error(void) { return g_err2.c_str(); }
static int fail2(int code, const std::string& msg) { g_err2 = msg; return code; }
#define CUDA_TRY2(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    return fail2(SDNS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

struct Space2 {
    int M[2];              // transform lengths
    AxisMap bmap0, fmap0;  // axis-0 maps: backward input (truncation / padding), forward output (3/2 truncation)
    int K1n, K1p;          // axis-1 modes entering the backward transform, even pitch
    double scale;          // 1 / (M0 M1)
};

struct sdns2d_plan {
    sdns2d_config cfg;
    int N[2], Nh, Nhp, prec, nf_in, nf_out, ncomp;
    size_t rs, cs;
    Space2 sp[2];
    cudaStream_t stream;
    std::vector<char> host_tables;
    std::map<int, size_t> tw_off;
    size_t kx_off, ky_off, off_tab, off_IN, off_W0, off_W2, off_S, ws_need;
    char* ws;
    long long launches;
};

static AxisMap all_map2(int n) { AxisMap m; m.nlo = n; m.nhi = 0; m.shift = 0; return m; }
static size_t align2(size_t x, size_t a) { return (x + a - 1) / a * a; }
static int default_kcut2(int n) { return (int)ceil(2.0 / 3.0 * (n / 2 + 1)) - 1; }      // spectralDNS3D_short.py:44-46

static bool size_ok2(int n) {
    switch (n) {
#define X(N) case N: return true;
        SDNS_SIZES(X)
#undef X
        default: return false;
    }
}

template <typename T>
static void fill_tables2(sdns2d_plan* p) {
    typedef typename C2<T>::type V;
    std::vector<char>& h = p->host_tables;
    auto add_tw = [&](int n) {
        if (p->tw_off.count(n)) return;
        size_t off = align2(h.size(), 256);
        h.resize(off + sizeof(V) * n);
        V* tw = reinterpret_cast<V*>(h.data() + off);
        for (int j = 0; j < n; ++j) {
            long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
            tw[j].x = (T)cosl(ang); tw[j].y = (T)sinl(ang);
        }
        p->tw_off[n] = off;
    };
    for (int s = 0; s < 2; ++s) for (int i = 0; i < 2; ++i) add_tw(p->sp[s].M[i]);
    auto add_k = [&](int n, int len, double L, bool real_axis) {
        size_t off = align2(h.size(), 256);
        h.resize(off + sizeof(T) * len);
        T* k = reinterpret_cast<T*>(h.data() + off);
        for (int i = 0; i < len; ++i) {
            int kk = real_axis ? i : (i < (n + 1) / 2 ? i : i - n);
            k[i] = (T)(((double)kk * 2.0 * M_PI) / L);          // k*2*pi/L in double, then cast (solvers/NS.py:38-41)
        }
        return off;
    };
    p->kx_off = add_k(p->N[0], p->N[0], p->cfg.L[0], false);
    p->ky_off = add_k(p->N[1], p->Nh, p->cfg.L[1], true);
}

extern "C" int sdns2d_plan_create(sdns2d_plan** out, const sdns2d_config* cfg) {
    if (!out || !cfg) return fail2(SDNS_ERR_ARG, "null argument");
    if (cfg->abi_version != SDNS_ABI_VERSION) return fail2(SDNS_ERR_ARG, "ABI version mismatch");
    if (cfg->precision != SDNS_SINGLE && cfg->precision != SDNS_DOUBLE) return fail2(SDNS_ERR_ARG, "precision");
    if (cfg->solver != SDNS_NS2D && cfg->solver != SDNS_BQ2D) return fail2(SDNS_ERR_ARG, "solver: SDNS_NS2D or SDNS_BQ2D");
    if (cfg->dealias < SDNS_DEALIAS_NONE || cfg->dealias > SDNS_DEALIAS_32) return fail2(SDNS_ERR_ARG, "dealias");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail2(SDNS_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (libsdns_b200 has no CPU fallback)");
    CUDA_TRY2(cudaSetDevice(cfg->device));
    sdns2d_plan* p = new sdns2d_plan();
    p->cfg = *cfg;
    for (int i = 0; i < 2; ++i) {
        p->N[i] = cfg->N[i];
        if (p->N[i] < 2 || p->N[i] % 2) { delete p; return fail2(SDNS_ERR_SIZE, "N must be even"); }
        if (!(cfg->L[i] > 0)) { delete p; return fail2(SDNS_ERR_ARG, "L must be positive"); }
    }
    p->Nh = p->N[1] / 2 + 1; p->Nhp = (p->Nh + 1) & ~1;
    p->prec = cfg->precision; p->rs = p->prec ? 8 : 4; p->cs = 2 * p->rs;
    p->stream = 0; p->ws = nullptr; p->launches = 0;
    p->ncomp = cfg->solver == SDNS_BQ2D ? 3 : 2;
    p->nf_in = p->ncomp + 1;                       // + curl_hat
    p->nf_out = cfg->solver == SDNS_BQ2D ? 4 : 2;
    Space2& t = p->sp[SDNS_SPACE_T];
    t.M[0] = p->N[0]; t.M[1] = p->N[1];
    t.bmap0 = all_map2(p->N[0]); t.fmap0 = all_map2(p->N[0]); t.K1n = p->Nh;
    Space2& d = p->sp[SDNS_SPACE_TP];
    d = t;
    if (cfg->dealias == SDNS_DEALIAS_23) {
        const int kc0 = cfg->kcut[0] >= 0 ? cfg->kcut[0] : default_kcut2(p->N[0]);
        if (2 * kc0 + 1 < p->N[0]) { d.bmap0.nlo = kc0 + 1; d.bmap0.nhi = kc0; d.bmap0.shift = 0; }
        const int kc1 = cfg->kcut[1] >= 0 ? cfg->kcut[1] : default_kcut2(p->N[1]);
        if (kc1 + 1 < p->Nh) d.K1n = kc1 + 1;
    } else if (cfg->dealias == SDNS_DEALIAS_32) {
        d.M[0] = 3 * p->N[0] / 2; d.M[1] = 3 * p->N[1] / 2;
        d.bmap0.nlo = p->N[0] / 2; d.bmap0.nhi = p->N[0] - p->N[0] / 2; d.bmap0.shift = d.M[0] - p->N[0];
        d.fmap0 = d.bmap0;
    }
    for (int s = 0; s < 2; ++s) {
        Space2& q = p->sp[s];
        q.K1p = (q.K1n + 1) & ~1;
        q.scale = 1.0 / ((double)q.M[0] * q.M[1]);
        for (int i = 0; i < 2; ++i)
            if (!size_ok2(q.M[i])) {
                char b[128]; snprintf(b, sizeof b, "no compiled transform of length %d (have 2^k and 3*2^k, 8..3072)", q.M[i]);
                delete p; return fail2(SDNS_ERR_SIZE, b);
            }
    }
    if (p->prec) fill_tables2<double>(p); else fill_tables2<float>(p);
    const size_t nfm = 4;
    size_t w0 = 0, w2 = 0;
    for (int s = 0; s < 2; ++s) {
        w0 = std::max(w0, nfm * p->sp[s].M[0] * (size_t)p->sp[s].K1p);
        w2 = std::max(w2, nfm * p->sp[s].M[0] * (size_t)p->Nhp);
    }
    p->off_tab = 0;
    p->off_IN = align2(p->host_tables.size(), 256);
    p->off_W0 = p->off_IN + align2(nfm * p->N[0] * (size_t)p->Nh * p->cs, 256);
    p->off_W2 = p->off_W0 + align2(w0 * p->cs, 256);
    p->off_S = p->off_W2 + align2(w2 * p->cs, 256);
    p->ws_need = p->off_S + align2(nfm * p->N[0] * (size_t)p->Nh * p->cs, 256);
    *out = p;
    return SDNS_OK;
}

extern "C" int sdns2d_plan_destroy(sdns2d_plan* p) { delete p; return SDNS_OK; }
extern "C" int sdns2d_workspace_bytes(const sdns2d_plan* p, size_t* bytes) {
    if (!p || !bytes) return fail2(SDNS_ERR_ARG, "null argument");
    *bytes = p->ws_need; return SDNS_OK;
}
extern "C" int sdns2d_plan_set_workspace(sdns2d_plan* p, void* dptr, size_t bytes) {
    if (!p || !dptr) return fail2(SDNS_ERR_ARG, "null argument");
    if (bytes < p->ws_need) return fail2(SDNS_ERR_WORKSPACE, "workspace too small");
    if ((uintptr_t)dptr % 256) return fail2(SDNS_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    p->ws = (char*)dptr;
    CUDA_TRY2(cudaMemcpyAsync(p->ws + p->off_tab, p->host_tables.data(), p->host_tables.size(), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY2(cudaStreamSynchronize(p->stream));
    return SDNS_OK;
}
extern "C" int sdns2d_plan_set_stream(sdns2d_plan* p, void* s) {
    if (!p) return fail2(SDNS_ERR_ARG, "null plan");
    p->stream = (cudaStream_t)s; return SDNS_OK;
}
extern "C" int sdns2d_sync(sdns2d_plan* p) {
    if (!p) return fail2(SDNS_ERR_ARG, "null plan");
    CUDA_TRY2(cudaStreamSynchronize(p->stream));
    return SDNS_OK;
}
extern "C" int sdns2d_shapes(const sdns2d_plan* p, int32_t sp[2], int32_t ph[2], int32_t pd[2]) {
    if (!p) return fail2(SDNS_ERR_ARG, "null plan");
    sp[0] = p->N[0]; sp[1] = p->Nh;
    for (int i = 0; i < 2; ++i) { ph[i] = p->sp[0].M[i]; pd[i] = p->sp[1].M[i]; }
    return SDNS_OK;
}
extern "C" int sdns2d_launch_count(const sdns2d_plan* p, long long* c) {
    if (!p || !c) return fail2(SDNS_ERR_ARG, "null argument");
    *c = p->launches; return SDNS_OK;
}

static int need_ws2(sdns2d_plan* p) {
    if (!p) return fail2(SDNS_ERR_ARG, "null plan");
    if (!p->ws) return fail2(SDNS_ERR_WORKSPACE, "call sdns2d_plan_set_workspace first");
    return SDNS_OK;
}
static int launch2(sdns2d_plan* p, int fam, int n, const void* args) {
    const int e = g_launch[fam][p->prec](n, args, p->stream);
    p->launches++;
    if (e == -1000) { char b[96]; snprintf(b, sizeof b, "no kernel for length %d (family %d)", n, fam); return fail2(SDNS_ERR_SIZE, b); }
    if (e == -1001) return fail2(SDNS_ERR_SIZE, "array too large for one pass");
    if (e != 0) { char b[160]; snprintf(b, sizeof b, "kernel launch (family %d, n=%d): %s", fam, n, cudaGetErrorString((cudaError_t)e)); return fail2(SDNS_ERR_CUDA, b); }
    return SDNS_OK;
}

// ---- the passes ---------------------------------------------------------------------------
template <typename T>
struct Pipe2 {
    typedef typename C2<T>::type V;
    sdns2d_plan* p;
    const Space2& q;
    Pipe2(sdns2d_plan* p_, int space) : p(p_), q(p_->sp[space]) {}
    V* buf(size_t off) const { return reinterpret_cast<V*>(p->ws + off); }
    const V* tw(int n) const { return reinterpret_cast<const V*>(p->ws + p->off_tab + p->tw_off.at(n)); }
    // B0: dense spectral (nf, N0, Nh) -> W0 (nf, M0, K1p), inverse c2c along axis 0 over the kept axis-1 modes
    int b0(const V* in, int nf) {
        StridedArgs<T> a; memset(&a, 0, sizeof a);
        a.in = in; a.out = buf(p->off_W0);
        a.in_fs = (long long)p->N[0] * p->Nh; a.in_ls = p->Nh; a.in_os = 0;
        a.out_fs = (long long)q.M[0] * q.K1p; a.out_ls = q.K1p; a.out_os = 0;
        a.cw = q.K1n; a.ncols = q.K1n; a.col_nlo = 1; a.col_gap = 0;
        a.imap = q.bmap0; a.omap = all_map2(q.M[0]);
        a.tw = tw(q.M[0]); a.scale = (T)1; a.nfields = nf; a.self = -1;
        return launch2(p, FAM_PLAIN_BWD, q.M[0], &a);
    }
    // F0: W2 (nf, M0, Nhp) -> dense spectral (nf, N0, Nh), forward c2c along axis 0 (+ 3/2 truncation)
    int f0(V* out, int nf) {
        StridedArgs<T> a; memset(&a, 0, sizeof a);
        a.in = buf(p->off_W2); a.out = out;
        a.in_fs = (long long)q.M[0] * p->Nhp; a.in_ls = p->Nhp; a.in_os = 0;
        a.out_fs = (long long)p->N[0] * p->Nh; a.out_ls = p->Nh; a.out_os = 0;
        a.cw = p->Nh; a.ncols = p->Nh; a.col_nlo = 1; a.col_gap = 0;
        a.imap = all_map2(q.M[0]); a.omap = q.fmap0;
        a.tw = tw(q.M[0]); a.scale = (T)1; a.nfields = nf; a.self = -1;
        return launch2(p, FAM_PLAIN_FWD, q.M[0], &a);
    }
    // Z: the contiguous axis.  in / out are W0 / W2 or the caller's real arrays.
    int z(int fam, const void* in, void* out, int nf, bool in_is_W0, bool out_is_W2) {
        ZArgs<T> a; memset(&a, 0, sizeof a);
        a.in = in; a.out = out;
        a.in_ls = in_is_W0 ? q.K1p : q.M[1]; a.in_fs = (long long)q.M[0] * a.in_ls;
        a.out_ls = out_is_W2 ? p->Nhp : q.M[1]; a.out_fs = (long long)q.M[0] * a.out_ls;
        a.nlines = q.M[0]; a.nin_keep = q.K1n; a.nout_keep = p->Nh; a.nf = nf;
        a.tw = tw(q.M[1]);
        a.scale = (fam == FAM_Z_C2R) ? (T)1 : (T)q.scale;
        return launch2(p, fam, q.M[1], &a);
    }
};

template <typename T>
static int backward2_t(sdns2d_plan* p, int space, int nc, const void* in, void* out) {
    typedef typename C2<T>::type V;
    Pipe2<T> P(p, space);
    for (int c0 = 0; c0 < nc; c0 += 4) {
        const int nf = std::min(4, nc - c0);
        int e;
        if ((e = P.b0(reinterpret_cast<const V*>(in) + (long long)c0 * p->N[0] * p->Nh, nf))) return e;
        if ((e = P.z(FAM_Z_C2R, P.buf(p->off_W0), reinterpret_cast<T*>(out) + (long long)c0 * P.q.M[0] * P.q.M[1], nf, true, false))) return e;
    }
    return SDNS_OK;
}
template <typename T>
static int forward2_t(sdns2d_plan* p, int space, int nc, const void* in, void* out) {
    typedef typename C2<T>::type V;
    Pipe2<T> P(p, space);
    for (int c0 = 0; c0 < nc; c0 += 4) {
        const int nf = std::min(4, nc - c0);
        int e;
        if ((e = P.z(FAM_Z_R2C, reinterpret_cast<const T*>(in) + (long long)c0 * P.q.M[0] * P.q.M[1], P.buf(p->off_W2), nf, false, true))) return e;
        if ((e = P.f0(reinterpret_cast<V*>(out) + (long long)c0 * p->N[0] * p->Nh, nf))) return e;
    }
    return SDNS_OK;
}
extern "C" int sdns2d_forward(sdns2d_plan* p, int space, int nc, const void* in, void* out) {
    int e = need_ws2(p); if (e) return e;
    if (space < 0 || space > 1 || nc < 1 || !in || !out) return fail2(SDNS_ERR_ARG, "sdns2d_forward: bad argument");
    return p->prec ? forward2_t<double>(p, space, nc, in, out) : forward2_t<float>(p, space, nc, in, out);
}
extern "C" int sdns2d_backward(sdns2d_plan* p, int space, int nc, const void* in, void* out) {
    int e = need_ws2(p); if (e) return e;
    if (space < 0 || space > 1 || nc < 1 || !in || !out) return fail2(SDNS_ERR_ARG, "sdns2d_backward: bad argument");
    return p->prec ? backward2_t<double>(p, space, nc, in, out) : backward2_t<float>(p, space, nc, in, out);
}

// ---- elementwise kernels ---------------------------------------------------------------------
// fields entering the backward transform: the state components and curl_hat = 1j*(K0 u1 - K1 u0)
// (cross2 of a 2-D field, maths/cross.py:30-35 with a scalar result; cython_maths.in:105-147)
template <typename T>
__global__ void prep2d_kernel(typename C2<T>::type* in, const typename C2<T>::type* u, int ncomp, const T* kx, const T* ky,
                              int N0, int Nh) {
    typedef typename C2<T>::type V;
    const long long n = (long long)N0 * Nh;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int i1 = (int)(i % Nh), i0 = (int)(i / Nh);
        const V u0 = u[i], u1 = u[n + i];
        in[i] = u0; in[n + i] = u1;
        if (ncomp == 3) in[2 * n + i] = u[2 * n + i];
        in[ncomp * n + i] = icross<T, V>(kx[i0], u1, ky[i1], u0);
    }
}
template <typename T>
__global__ void cross2_2d_kernel(typename C2<T>::type* c, const typename C2<T>::type* u, const T* kx, const T* ky, int N0, int Nh) {
    typedef typename C2<T>::type V;
    const long long n = (long long)N0 * Nh;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int i1 = (int)(i % Nh), i0 = (int)(i / Nh);
        c[i] = icross<T, V>(kx[i0], u[n + i], ky[i1], u[i]);
    }
}

struct Epi2Args {
    const void* conv;      // (nf_out, N0, Nh): transformed products (NULL: du already holds the convection, stand-alone call)
    const void* u_hat; void* rhs; void* u0; void* u1; void* u2; void* p_hat; const void* source;
    double nu, Ri, Pr, adt, bdt;
    int solver, mask_nyquist, out_mode, rk, N0, N1, Nh;
};
// Nyquist mask (NS.py:253-254), pressure + diffusion (NS2D: cython_solvers.in:105-127; Bq2D: Bq2D.py:139-156,
// cython_solvers.in:82-103), Source (NS.py:259), and the RK4 stage update (maths/integrators.py:150-159)
template <typename T>
__global__ void epi2d_kernel(const Epi2Args a, const T* kx, const T* ky) {
    typedef typename C2<T>::type V;
    const long long n = (long long)a.N0 * a.Nh;
    const V* conv = reinterpret_cast<const V*>(a.conv);
    const V* uh = reinterpret_cast<const V*>(a.u_hat);
    V* rhs = reinterpret_cast<V*>(a.rhs);
    const int nc = a.solver == SDNS_BQ2D ? 3 : 2;
    const T nu = (T)a.nu, Ri = (T)a.Ri, Pr = (T)a.Pr, adt = (T)a.adt, bdt = (T)a.bdt;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int i1 = (int)(i % a.Nh), i0 = (int)(i / a.Nh);
        const T k0 = kx[i0], k1 = ky[i1];
        T ksq = k0 * k0; ksq += k1 * k1;
        const T ks = ksq == (T)0 ? (T)1 : ksq;
        const T q0 = k0 / ks, q1 = k1 / ks;
        V d[3];
        if (conv) {
            d[0] = conv[i]; d[1] = conv[n + i];
            if (nc == 3) {
                // rhs[2] = -1j*(K0 F0 + K1 F1), F = F(u*rho)   (Bq2D.py:135-137)
                const V f0 = conv[2 * n + i], f1 = conv[3 * n + i];
                const T sx = k0 * f0.x + k1 * f1.x, sy = k0 * f0.y + k1 * f1.y;
                d[2].x = sy; d[2].y = -sx;
            }
            if (a.mask_nyquist && (2 * i0 == a.N0 || 2 * i1 == a.N1)) { d[0] = V(); d[1] = V(); d[2] = V(); }
        } else {
            d[0] = rhs[i]; d[1] = rhs[n + i]; if (nc == 3) d[2] = rhs[2 * n + i];
        }
        const V w0 = uh[i], w1 = uh[n + i];
        V w2 = V(); if (nc == 3) w2 = uh[2 * n + i];
        const T z = nu * ksq;
        V ph;
        ph.x = d[0].x * q0 + d[1].x * q1; ph.y = d[0].y * q0 + d[1].y * q1;
        if (nc == 3) { ph.x -= Ri * w2.x * q1; ph.y -= Ri * w2.y * q1; }
        if (a.p_hat) reinterpret_cast<V*>(a.p_hat)[i] = ph;
        d[0].x -= ph.x * k0 + w0.x * z; d[0].y -= ph.y * k0 + w0.y * z;
        if (nc == 3) {
            d[1].x -= ph.x * k1 + w1.x * z + Ri * w2.x; d[1].y -= ph.y * k1 + w1.y * z + Ri * w2.y;
            d[2].x -= w2.x * z / Pr; d[2].y -= w2.y * z / Pr;
        } else {
            d[1].x -= ph.x * k1 + w1.x * z; d[1].y -= ph.y * k1 + w1.y * z;
        }
        if (a.source) {
            const V* s = reinterpret_cast<const V*>(a.source);
            for (int c = 0; c < nc; ++c) { d[c].x += s[c * n + i].x; d[c].y += s[c * n + i].y; }
        }
        const V w[3] = {w0, w1, w2};
        if (a.out_mode == OUT_RHS) {
            for (int c = 0; c < nc; ++c) rhs[c * n + i] = d[c];
        } else {
            V* u0 = reinterpret_cast<V*>(a.u0); V* u1 = reinterpret_cast<V*>(a.u1); V* u2 = reinterpret_cast<V*>(a.u2);
            for (int c = 0; c < nc; ++c) {
                const long long o = c * n + i;
                V b1, b2;
                if (a.rk == 0) { b1 = w[c]; b2 = w[c]; u1[o] = b1; }
                else { b2 = u2[o]; b1 = u1[o]; }
                b2.x += adt * d[c].x; b2.y += adt * d[c].y;
                if (a.rk < 3) {
                    u2[o] = b2;
                    V nw; nw.x = b1.x + bdt * d[c].x; nw.y = b1.y + bdt * d[c].y;
                    u0[o] = nw;
                } else {
                    u0[o] = b2;
                }
            }
        }
    }
}

template <typename T>
static int epilogue2(sdns2d_plan* p, const Epi2Args& a) {
    SDNS_LAUNCH(epi2d_kernel<T>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)(a, (const T*)(p->ws + p->kx_off), (const T*)(p->ws + p->ky_off));
    p->launches++;
    CUDA_TRY2(cudaGetLastError());
    return SDNS_OK;
}

template <typename T>
static int rhs2_t(sdns2d_plan* p, const void* u_hat, Epi2Args ea) {
    typedef typename C2<T>::type V;
    Pipe2<T> P(p, SDNS_SPACE_TP);
    int e;
    SDNS_LAUNCH(prep2d_kernel<T>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)(P.buf(p->off_IN), reinterpret_cast<const V*>(u_hat), p->ncomp,
        (const T*)(p->ws + p->kx_off), (const T*)(p->ws + p->ky_off), p->N[0], p->Nh);
    p->launches++;
    CUDA_TRY2(cudaGetLastError());
    if ((e = P.b0(P.buf(p->off_IN), p->nf_in))) return e;
    if ((e = P.z(p->cfg.solver == SDNS_BQ2D ? FAM_Z_BQ2D : FAM_Z_NS2D, P.buf(p->off_W0), P.buf(p->off_W2), p->nf_in, true, true))) return e;
    if ((e = P.f0(P.buf(p->off_S), p->nf_out))) return e;
    ea.conv = P.buf(p->off_S); ea.u_hat = u_hat;
    ea.solver = p->cfg.solver; ea.mask_nyquist = p->cfg.mask_nyquist; ea.N0 = p->N[0]; ea.N1 = p->N[1]; ea.Nh = p->Nh;
    return epilogue2<T>(p, ea);
}

extern "C" int sdns2d_compute_rhs(sdns2d_plan* p, void* rhs, const void* u_hat, double nu, double Ri, double Pr,
                                  const void* source, void* p_hat) {
    int e = need_ws2(p); if (e) return e;
    if (!rhs || !u_hat) return fail2(SDNS_ERR_ARG, "sdns2d_compute_rhs: null array");
    Epi2Args a; memset(&a, 0, sizeof a);
    a.rhs = rhs; a.p_hat = p_hat; a.source = source; a.nu = nu; a.Ri = Ri; a.Pr = Pr; a.out_mode = OUT_RHS;
    return p->prec ? rhs2_t<double>(p, u_hat, a) : rhs2_t<float>(p, u_hat, a);
}

extern "C" int sdns2d_rk4_step(sdns2d_plan* p, void* u_hat, void* u1, void* u2, double dt, double nu, double Ri, double Pr,
                               const void* source) {
    int e = need_ws2(p); if (e) return e;
    if (!u_hat || !u1 || !u2) return fail2(SDNS_ERR_ARG, "sdns2d_rk4_step: null array");
    for (int rk = 0; rk < 4; ++rk) {
        Epi2Args a; memset(&a, 0, sizeof a);
        a.out_mode = OUT_STAGE; a.u0 = u_hat; a.u1 = u1; a.u2 = u2; a.source = source; a.rk = rk; a.nu = nu; a.Ri = Ri; a.Pr = Pr;
        // a, b of maths/integrators.py:185-186 in context.float, products with dt in that type
        if (p->prec) {
            const double ca[4] = {1. / 6., 1. / 3., 1. / 3., 1. / 6.}, cb[3] = {0.5, 0.5, 1.};
            a.adt = ca[rk] * dt; a.bdt = rk < 3 ? cb[rk] * dt : 0.0;
        } else {
            const float ca[4] = {(float)(1. / 6.), (float)(1. / 3.), (float)(1. / 3.), (float)(1. / 6.)}, cb[3] = {0.5f, 0.5f, 1.f};
            a.adt = (double)(float)(ca[rk] * (float)dt); a.bdt = rk < 3 ? (double)(float)(cb[rk] * (float)dt) : 0.0;
        }
        e = p->prec ? rhs2_t<double>(p, u_hat, a) : rhs2_t<float>(p, u_hat, a);
        if (e) return e;
    }
    return SDNS_OK;
}

// ForwardEuler / AB2 (maths/integrators.py:161-175) on the 2-D state
template <typename T>
__global__ void euler2d_kernel(typename C2<T>::type* u, const typename C2<T>::type* r, T dt, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        typename C2<T>::type a = u[i], b = r[i];
        a.x += b.x * dt; a.y += b.y * dt; u[i] = a;
    }
}
template <typename T>
__global__ void ab2_2d_kernel(typename C2<T>::type* u, typename C2<T>::type* u1, const typename C2<T>::type* r, T dt, int first, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        typename C2<T>::type a = u[i], b = r[i], o = u1[i];
        const T rx = b.x * dt, ry = b.y * dt;
        if (first) { a.x += rx; a.y += ry; }
        else { a.x += ((T)1.5 * rx - (T)0.5 * o.x); a.y += ((T)1.5 * ry - (T)0.5 * o.y); }
        u[i] = a; o.x = rx; o.y = ry; u1[i] = o;
    }
}
extern "C" int sdns2d_euler_step(sdns2d_plan* p, void* u_hat, void* rhs, double dt, double nu, double Ri, double Pr, const void* source) {
    int e = sdns2d_compute_rhs(p, rhs, u_hat, nu, Ri, Pr, source, nullptr); if (e) return e;
    const long long n = (long long)p->ncomp * p->N[0] * p->Nh;
    if (p->prec) SDNS_LAUNCH(euler2d_kernel<double>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)((double2*)u_hat, (const double2*)rhs, dt, n);
    else SDNS_LAUNCH(euler2d_kernel<float>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)((float2*)u_hat, (const float2*)rhs, (float)dt, n);
    p->launches++;
    CUDA_TRY2(cudaGetLastError());
    return SDNS_OK;
}
extern "C" int sdns2d_ab2_step(sdns2d_plan* p, void* u_hat, void* u1, void* rhs, double dt, int tstep, double nu, double Ri, double Pr,
                               const void* source) {
    int e = sdns2d_compute_rhs(p, rhs, u_hat, nu, Ri, Pr, source, nullptr); if (e) return e;
    if (!u1) return fail2(SDNS_ERR_ARG, "sdns2d_ab2_step: null array");
    const long long n = (long long)p->ncomp * p->N[0] * p->Nh;
    if (p->prec) SDNS_LAUNCH(ab2_2d_kernel<double>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)((double2*)u_hat, (double2*)u1, (const double2*)rhs, dt, tstep == 0, n);
    else SDNS_LAUNCH(ab2_2d_kernel<float>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)((float2*)u_hat, (float2*)u1, (const float2*)rhs, (float)dt, tstep == 0, n);
    p->launches++;
    CUDA_TRY2(cudaGetLastError());
    return SDNS_OK;
}

// stand-alone operators of the fine-grained plug-in surface (optimization/__init__.py:12-55)
extern "C" int sdns2d_cross2(sdns2d_plan* p, void* c, const void* u_hat) {
    int e = need_ws2(p); if (e) return e;
    if (!c || !u_hat || c == u_hat) return fail2(SDNS_ERR_ARG, "sdns2d_cross2: c and u_hat must be distinct arrays");
    if (p->prec) SDNS_LAUNCH(cross2_2d_kernel<double>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)((double2*)c, (const double2*)u_hat,
        (const double*)(p->ws + p->kx_off), (const double*)(p->ws + p->ky_off), p->N[0], p->Nh);
    else SDNS_LAUNCH(cross2_2d_kernel<float>, SDNS2D_EW_BLOCKS, 256, 0, p->stream)((float2*)c, (const float2*)u_hat,
        (const float*)(p->ws + p->kx_off), (const float*)(p->ws + p->ky_off), p->N[0], p->Nh);
    p->launches++;
    CUDA_TRY2(cudaGetLastError());
    return SDNS_OK;
}
extern "C" int sdns2d_add_pressure_diffusion(sdns2d_plan* p, void* du, const void* u_hat, double nu, double Ri, double Pr, void* p_hat) {
    int e = need_ws2(p); if (e) return e;
    if (!du || !u_hat || du == u_hat) return fail2(SDNS_ERR_ARG, "sdns2d_add_pressure_diffusion: du and u_hat must be distinct arrays");
    Epi2Args a; memset(&a, 0, sizeof a);
    a.conv = nullptr; a.u_hat = u_hat; a.rhs = du; a.p_hat = p_hat; a.nu = nu; a.Ri = Ri; a.Pr = Pr; a.out_mode = OUT_RHS;
    a.solver = p->cfg.solver; a.mask_nyquist = 0; a.N0 = p->N[0]; a.N1 = p->N[1]; a.Nh = p->Nh;
    return p->prec ? epilogue2<double>(p, a) : epilogue2<float>(p, a);
}
