/* sdns_b200.h -- C ABI of libsdns_b200.so: the B200-native replacement for the hot path of
 * spectralDNS (triply periodic pseudo-spectral RHS + RK4 step of the NS / VV / MHD solvers).
 *
 * The reference has no FFI of its own for this path: its "native" layer is Cython modules
 * called with numpy arrays (spectralDNS/optimization/cython_{maths,solvers,integrators}.in) plus
 * the FFTW/MPI calls that shenfun / mpi4py-fft make underneath T.forward / T.backward.  Each
 * entry point below names the reference interface it replaces (paths relative to the reference
 * root).  INTEGRATION.md shows the ctypes binding a maintainer adds to solvers/NS.py etc.
 *
 * Conventions
 *   - plain C linkage, plain pointers and sizes; no C++/torch types cross the boundary;
 *   - every function returns 0 on success or a negative sdns_status; sdns_last_error() returns
 *     a thread-local message for the last failure;
 *   - device pointers are BORROWED (the caller -- PyTorch in the shipped host layer -- owns
 *     the allocations and keeps them alive); a plan is not thread-safe;
 *   - all work is enqueued on the plan's CUDA stream (sdns_plan_set_stream; default stream 0)
 *     and is asynchronous unless stated otherwise;
 *   - spectral arrays: C order (ncomp, N0, N1/P, N2/2+1) complex<real>, physical arrays
 *     (ncomp, M0/P, M1, M2) real, exactly the local shapes of the reference's Function(VT) /
 *     Array(VT) (solvers/NS.py:51-63, spectralDNS3D_short.py:28-29).
 *   - there is NO CPU fallback: every entry point fails with SDNS_ERR_CUDA when no device is
 *     present.
 */
#ifndef SDNS_B200_H
#define SDNS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDNS_ABI_VERSION 1

typedef enum {
    SDNS_OK = 0,
    SDNS_ERR_ARG = -1,        /* invalid argument / unsupported configuration */
    SDNS_ERR_SIZE = -2,       /* grid size without a compiled transform */
    SDNS_ERR_CUDA = -3,       /* CUDA runtime error (message has the cudaError string) */
    SDNS_ERR_WORKSPACE = -4,  /* workspace missing or too small */
    SDNS_ERR_STATE = -5       /* call order violation */
} sdns_status;

enum { SDNS_SINGLE = 0, SDNS_DOUBLE = 1 };                         /* config.py:166-167  --precision   */
enum { SDNS_DEALIAS_NONE = 0, SDNS_DEALIAS_23 = 1, SDNS_DEALIAS_32 = 2 }; /* config.py:190-192 --dealias */
enum { SDNS_NS = 0, SDNS_VV = 1, SDNS_MHD = 2 };                   /* config.py:228-233 solver sub-command */
enum { SDNS_CONV_VORTEX = 0, SDNS_CONV_DIVERGENCE = 1,
       SDNS_CONV_STANDARD = 2, SDNS_CONV_SKEWED = 3 };             /* config.py:214-216 --convection   */
enum { SDNS_SLAB = 0, SDNS_PENCIL = 1 };                           /* config.py:194-195 --decomposition */
/* Slab ownership of the spectral axis 1.  BLOCKS is the reference's layout: rank r owns k1 in
 * [r*N1/P, (r+1)*N1/P) (solvers/spectralinit.py:19-21, spectralDNS3D_short.py:29), i.e. the slice
 * [r*N1l : (r+1)*N1l].  CYCLIC deals the modes out round robin, rank r owns the slice [r::P]: under the
 * 2/3 rule every rank then holds the same number of kept modes (with BLOCKS the ranks at the ends of the
 * k1 range hold 1.5x the average and the middle ones none).  Same global field either way;
 * sdns_k1_layout reports the slice. */
enum { SDNS_K1_BLOCKS = 0, SDNS_K1_CYCLIC = 1 };
enum { SDNS_SPACE_T = 0, SDNS_SPACE_TP = 1 };                      /* solvers/NS.py:21-32  T/VT vs Tp/VTp */

typedef struct sdns_config {
    int32_t abi_version;      /* SDNS_ABI_VERSION */
    int32_t N[3];             /* params.N  (config.py:115-118) */
    double  L[3];             /* params.L  (config.py:141-145) */
    int32_t precision;        /* SDNS_SINGLE | SDNS_DOUBLE */
    int32_t dealias;          /* SDNS_DEALIAS_* (solvers/NS.py:29-31) */
    int32_t solver;           /* SDNS_NS | SDNS_VV | SDNS_MHD */
    int32_t convection;       /* SDNS_CONV_* (NS.getConvection, solvers/NS.py:164-201) */
    int32_t mask_nyquist;     /* params.mask_nyquist (config.py:207-209, NS.py:34) */
    int32_t decomposition;    /* SDNS_SLAB | SDNS_PENCIL */
    int32_t kcut[3];          /* 2/3-rule: largest kept |k| per axis; <0 = default
                                 ceil(2/3*(N/2+1))-1  (spectralDNS3D_short.py:44-46) */
    int32_t prune;            /* 1: skip transform lines that the truncation zeroes (same result) */
    int32_t rank, nranks;     /* slab position of this process (solvers/spectralinit.py:19-21); an extent the rank count
                                 does not divide is split N / P per rank, the first N % P ranks one more (mpi4py-fft) */
    int32_t device;           /* CUDA device ordinal */
    int32_t k1_layout;        /* SDNS_K1_BLOCKS | SDNS_K1_CYCLIC: which axis-1 modes a rank owns (nranks > 1) */
    int32_t reserved[7];
} sdns_config;

typedef struct sdns_plan sdns_plan;

/* version / diagnostics */
int         sdns_abi_version(void);
const char* sdns_last_error(void);
/* 0 if a transform of length n is compiled in for the precision, else SDNS_ERR_SIZE */
int         sdns_size_supported(int n, int precision);

/* plan lifetime: replaces get_context()'s FunctionSpace/TensorProductSpace/get_dealiased setup
 * (solvers/NS.py:12-48, MHD.py:13-50): spaces, wavenumbers K, K2, K_over_K2, Nyquist mask. */
int sdns_plan_create(sdns_plan** plan, const sdns_config* cfg);
int sdns_plan_destroy(sdns_plan* plan);
/* scratch the caller must provide (device bytes), then hand over with set_workspace.  Replaces
 * the reference's work = CachedArrayDict() arrays and u_dealias / ZZ_hat (NS.py:57-64, MHD.py:59-60) */
int sdns_workspace_bytes(const sdns_plan* plan, size_t* bytes);
int sdns_plan_set_workspace(sdns_plan* plan, void* device_ptr, size_t bytes);
int sdns_plan_set_stream(sdns_plan* plan, void* cuda_stream);
int sdns_sync(sdns_plan* plan);                          /* cudaStreamSynchronize on the plan stream */

/* Slab decomposition, one process per GPU of one node (replaces mpi4py-fft's Pencil/Transfer =
 * MPI_Alltoallw inside every shenfun transform; in-tree analogue spectralDNS3D_short.py:50-62).
 * For nranks > 1 the library owns the workspace (cudaMalloc) so that it can be mapped into the
 * peers with CUDA IPC: every rank calls sdns_comm_alloc, publishes the 64-byte handle from
 * sdns_comm_handle to all ranks (the host layer uses torch.distributed.all_gather), then calls
 * sdns_comm_open with the nranks handles in rank order.  The transposes have no kernel of their
 * own: the pass in front of each one stores directly into the owning GPU's buffer over NVLink and a
 * device-side flag barrier orders the passes.  sdns_comm_status reports a barrier timeout. */
int sdns_comm_alloc(sdns_plan* plan);
int sdns_comm_handle(sdns_plan* plan, void* handle64);
int sdns_comm_open(sdns_plan* plan, const void* handles, int nranks);
int sdns_comm_status(sdns_plan* plan, int* timed_out);

/* local array extents of this rank (T.shape(True), T.shape(False), Tp.shape(False)) */
int sdns_local_shapes(const sdns_plan* plan, int32_t spectral[3], int32_t physical[3], int32_t padded[3]);
/* axis-1 mode index of local spectral index j: first + j*step (the local_slice of T.local_slice(True)[1]) */
int sdns_k1_layout(const sdns_plan* plan, int32_t* first, int32_t* step);

/* T.forward / VT.forward (space = SDNS_SPACE_T) and Tp.forward / VTp.forward (SDNS_SPACE_TP):
 * rfftn/prod(M) incl. 3/2-rule truncation.  Call sites NS.py:103,135; MHD.py:107. */
int sdns_forward(sdns_plan* plan, int space, int ncomp, const void* real_in, void* cplx_out);
/* T.backward / VT.backward / Tp.backward / VTp.backward: irfftn*prod(M) incl. the 2/3-rule
 * truncation of the input or the 3/2-rule zero padding.  Call sites NS.py:93,98,128; MHD.py:121. */
int sdns_backward(sdns_plan* plan, int space, int ncomp, const void* cplx_in, void* real_out);

/* solver.ComputeRHS(rhs, u_hat, solver, **context)  (NS.py:219-261, VV.py:112-146, MHD.py:151-176)
 * including conv (getConvection), mask_nyquist, add_pressure_diffusion / add_linear, +Source.
 * source and p_hat may be NULL.  nu/eta as in params (config.py:123-127: cast to the precision). */
int sdns_compute_rhs(sdns_plan* plan, void* rhs, const void* u_hat, double nu, double eta,
                     const void* source, void* p_hat);

/* solver.conv(rhs, u_hat, ...) alone, i.e. the function returned by getConvection()
 * (NS.py:164-201, VV.py:85-103, MHD.py:112-130): the dealiased nonlinear term without the Nyquist
 * mask, pressure and diffusion. */
int sdns_compute_conv(sdns_plan* plan, void* rhs, const void* u_hat);

/* integrate() for params.integrator == 'RK4'  (maths/integrators.py:150-159,177-191;
 * cython_integrators.in:8-52): four ComputeRHS evaluations with the stage updates fused into the
 * last transform pass.  u_hat is updated in place; u1, u2 are the integrator's work arrays
 * (u0.copy() at integrators.py:181,187): same size as u_hat, contents unspecified on return (the
 * library keeps them in a k1-major layout internally). */
int sdns_rk4_step(sdns_plan* plan, void* u_hat, void* u1, void* u2, double dt, double nu, double eta,
                  const void* source);

/* ForwardEuler and AB2 (maths/integrators.py:161-175): rhs is the caller's dU array. */
int sdns_euler_step(sdns_plan* plan, void* u_hat, void* rhs, double dt, double nu, double eta,
                    const void* source);
int sdns_ab2_step(sdns_plan* plan, void* u_hat, void* u1, void* rhs, double dt, int tstep,
                  double nu, double eta, const void* source);

/* cross2(c, K, b) / cross2(c, K_over_K2, b): c = 1j*(K x b)  (maths/cross.py:30-35,
 * cython_maths.in:32-86).  over_k2 != 0 selects K_over_K2 (VV.py:64). */
int sdns_cross2(sdns_plan* plan, void* c, const void* b, int over_k2);

/* cross1(c, a, b): real c = a x b over n points per component (maths/cross.py:16-28,
 * cython_maths.in:13-30); cross2 with a dense real a (cython_maths.in:39-60); project(u, K, K_over_K2)
 * (maths/maths.py:8-11).  Stand-alone versions of operators that the RHS kernels fuse. */
int sdns_cross1(sdns_plan* plan, void* c, const void* a, const void* b, long long n);
int sdns_cross2_dense(sdns_plan* plan, void* c, const void* a_real, const void* b);
int sdns_project(sdns_plan* plan, void* u_hat);
/* add_pressure_diffusion_NS(du, u_hat, nu, ksq, kk, p_hat, k_over_k2)  (solvers/NS.py:203-217; compiled form
 * optimization/cython_solvers.in:40-80, dispatched at optimization/__init__.py:12-55): in place
 * p_hat = sum_i du_i K_i/K^2, du_i -= p_hat K_i + nu K^2 u_hat_i.  p_hat (one spectral component) may be NULL.
 * Stand-alone version of what the F0 pass fuses into sdns_compute_rhs. */
int sdns_add_pressure_diffusion(sdns_plan* plan, void* du, const void* u_hat, double nu, void* p_hat);

/* Building blocks of adaptiveRK (BS5_adaptive / BS5_fixed, maths/integrators.py:15-147,193-225):
 * out = base + sum_t coeffs[t]*arrays[t] over ncomp spectral components (base may be NULL), and the
 * per-component error sums sum |err/(atol + max(|u0|,|u1|)*rtol)|^2 of integrators.py:86-92
 * (synchronous; out[ncomp], local block -- the caller reduces over ranks). */
int sdns_lincomb(sdns_plan* plan, void* out, const void* base, int nterms, const double* coeffs,
                 const void* const* arrays, int ncomp);
int sdns_errnorm(sdns_plan* plan, const void* u0, const void* u1, const void* err, double atol, double rtol,
                 int ncomp, double* out);

/* shenfun.fourier.energy_fourier(u_hat, T) of ncomp components (tests/TG.py:101,
 * demo/Isotropic.py:67,167-182): Hermitian-weighted sum |u_hat|^2 of the LOCAL block.
 * Synchronous (returns the value). */
int sdns_energy(sdns_plan* plan, const void* u_hat, int ncomp, double* out);

/* Diagnostics and low-wavenumber forcing of demo/Isotropic.py, computed where the state lives (the reference does
 * them with numpy on the context's host arrays).  All reductions cover the LOCAL block and are synchronous; the
 * caller sums over ranks.
 *   sdns_energy_weighted : energy_fourier(U_hat*weight, T), weight a real field of the spectral shape or NULL
 *                          (demo/Isotropic.py:167-169; weight_is_double selects float64 / float32 storage)
 *   sdns_scale_field     : U_hat *= a*factor + b*(1 - factor), factor a real field broadcast over the components:
 *                          (a, b) = (1, 0) multiplies by the field; (alpha, 1) with factor = k2_mask is the forcing
 *                          rescale U_hat *= alpha*k2_mask + (1 - k2_mask) of Isotropic.py:180
 *   sdns_set_mode        : U_hat[:, i0, i1, i2] = re + i im, i1 local (Isotropic.py:63-64, 162-163: the mean mode)
 *   sdns_enstrophy       : energy_fourier(cross2(K, U_hat), T), the `dissipation` of Isotropic.py:243-244
 *   sdns_divergence_norm : sum w |i K.U_hat|^2 = L2_norm(get_divergence(...)) by Parseval (Isotropic.py:245-247)
 *   sdns_spectrum        : shell sums and point counts of spectrum() (Isotropic.py:88-118): shell i holds the modes
 *                          with i + 0.5 < |K| <= i + 1.5; sums[i] = sum w sum_c |U_hat_c|^2 (w = 1 on the first and
 *                          last k2 plane, else 2), counts[i] = number of modes; 2 <= nbins <= 4096, shell nbins-1 unused */
int sdns_energy_weighted(sdns_plan* plan, const void* u_hat, int ncomp, const void* weight, int weight_is_double, double* out);
int sdns_scale_field(sdns_plan* plan, void* u_hat, int ncomp, const void* factor, int factor_is_double, double a, double b);
int sdns_set_mode(sdns_plan* plan, void* u_hat, int ncomp, int i0, int i1, int i2, double re, double im);
int sdns_enstrophy(sdns_plan* plan, const void* u_hat, double* out);
int sdns_divergence_norm(sdns_plan* plan, const void* u_hat, double* out);
int sdns_spectrum(sdns_plan* plan, const void* u_hat, int ncomp, int nbins, double* sums, double* counts);

/* Host-buffer variant of integrate(): copies the state from (pinned) host memory, runs nsteps RK4
 * steps, copies it back.  This is the call a host-array caller (the reference's numpy context)
 * makes; bench.py's e2e number is measured through it. */
int sdns_rk4_steps_host(sdns_plan* plan, void* host_u_hat, void* dev_u_hat, void* dev_u1, void* dev_u2,
                        int nsteps, double dt, double nu, double eta);

/* number of kernels this plan has launched since creation (bench.py's gpu_launches) */
int sdns_launch_count(const sdns_plan* plan, long long* count);

/* Measurement aid (no reference counterpart; the reference's own metric is Timer, utilities/__init__.py:18-68):
 * with profiling on, every kernel launch is bracketed by CUDA events on the plan stream.
 * sdns_profile_read returns, per kernel family (0 plain fwd c2c, 1 plain bwd c2c, 2 NS B0, 3 VV B0,
 * 4 NS F0, 5 VV F0, 6 MHD F0, 7 c2r, 8 r2c, 9 fused z cross, 10 fused z MHD, 11 NS gradient B0,
 * 12 z dot product, 13 z symmetric products, 14 NS divergence-form F0), the summed device
 * time, the launch count and the summed algorithmic HBM bytes since sdns_profile_enable. */
int sdns_profile_enable(sdns_plan* plan, int on);
int sdns_profile_read(sdns_plan* plan, int family, double* total_ms, long long* launches, double* bytes);
/* bytes the family's launches stored into peer GPUs over NVLink (slab transposes) */
int sdns_profile_read_nvlink(sdns_plan* plan, int family, double* bytes);
/* copy-engine exchange (the default for nranks > 1; SDNS_EXCHANGE=store selects the fused peer stores): bytes this
 * rank sent over NVLink since sdns_profile_enable, the number of strided copies, and the busy time of the busiest
 * per-peer copy stream (the copies to different peers run concurrently) */
int sdns_profile_read_copies(sdns_plan* plan, double* busy_ms, double* bytes, long long* ncopies);
/* sdns_profile_enable(plan, 2) additionally keeps a timeline: start and end (ms after the enable call, device clock)
 * of every kernel launch, peer copy and cross-GPU barrier.  rows receives up to max_rows records of four doubles
 * (kind, t_start_ms, t_end_ms, bytes); kind = kernel family 0..14, 99 = barrier, 100 + s = copy on copy stream s.
 * *nrows is the number of records available (rows may be NULL to query it).  sdns_profile_enable clears it. */
int sdns_profile_timeline(sdns_plan* plan, double* rows, int max_rows, int* nrows);
/* transfer-role exchange (the default for nranks > 1, csrc/xfer.cuh; replaces mpi4py-fft's Alltoallw Transfer, in-tree
 * analogue spectralDNS3D_short.py:50-62): bytes this rank has sent over NVLink since plan creation, and how many
 * transfer-only launches were needed for what no pass kernel could carry.  In the timeline those launches are
 * kind 98. */
int sdns_xfer_stats(sdns_plan* plan, double* bytes, long long* flush_launches);

/* ---- doubly periodic (2-D) solvers: NS2D (solvers/NS2D.py:13-51) and Bq2D (solvers/Bq2D.py:101-176) -----------------
 * Same conventions as above.  Spectral arrays are C order (ncomp, N0, N1/2+1) complex, physical arrays (ncomp, M0, M1)
 * real, ncomp = 2 (NS2D: u) or 3 (Bq2D: u, rho) -- the shapes of Function(VT) / Function(VM) in the reference
 * (NS.py:51-52 in two dimensions, Bq2D.py:52-63).  Single GPU. */
enum { SDNS_NS2D = 0, SDNS_BQ2D = 1 };                             /* config.py:256-261 doublyperiodic sub-commands */
typedef struct sdns2d_config {
    int32_t abi_version;      /* SDNS_ABI_VERSION */
    int32_t N[2];             /* params.N */
    double  L[2];             /* params.L  (config.py:246) */
    int32_t precision;        /* SDNS_SINGLE | SDNS_DOUBLE */
    int32_t dealias;          /* SDNS_DEALIAS_* */
    int32_t solver;           /* SDNS_NS2D | SDNS_BQ2D */
    int32_t mask_nyquist;     /* params.mask_nyquist */
    int32_t kcut[2];          /* 2/3-rule cutoff per axis, <0 = default (see sdns_config.kcut) */
    int32_t device;
    int32_t reserved[8];
} sdns2d_config;
typedef struct sdns2d_plan sdns2d_plan;
const char* sdns2d_last_error(void);
/* get_context() of NS2D / Bq2D: spaces T, Tp, wavenumbers (NS2D.py:13-18, Bq2D.py:13-50) */
int sdns2d_plan_create(sdns2d_plan** plan, const sdns2d_config* cfg);
int sdns2d_plan_destroy(sdns2d_plan* plan);
int sdns2d_workspace_bytes(const sdns2d_plan* plan, size_t* bytes);
int sdns2d_plan_set_workspace(sdns2d_plan* plan, void* device_ptr, size_t bytes);
int sdns2d_plan_set_stream(sdns2d_plan* plan, void* cuda_stream);
int sdns2d_sync(sdns2d_plan* plan);
int sdns2d_shapes(const sdns2d_plan* plan, int32_t spectral[2], int32_t physical[2], int32_t padded[2]);
int sdns2d_launch_count(const sdns2d_plan* plan, long long* count);
/* T.forward / T.backward and Tp.forward / Tp.backward of ncomp fields (NS2D.py:20-31, 43-47; tests/TG2D.py:12-16) */
int sdns2d_forward(sdns2d_plan* plan, int space, int ncomp, const void* real_in, void* cplx_out);
int sdns2d_backward(sdns2d_plan* plan, int space, int ncomp, const void* cplx_in, void* real_out);
/* ComputeRHS (NS.py:219-261 with NS2D's Conv, NS2D.py:33-51; Bq2D.py:158-186): conv, Nyquist mask,
 * add_pressure_diffusion, + Source (NS2D).  Ri and Pr are ignored by NS2D.  source and p_hat may be NULL. */
int sdns2d_compute_rhs(sdns2d_plan* plan, void* rhs, const void* u_hat, double nu, double Ri, double Pr,
                       const void* source, void* p_hat);
/* integrate() for RK4 / ForwardEuler / AB2 (maths/integrators.py:150-175) */
int sdns2d_rk4_step(sdns2d_plan* plan, void* u_hat, void* u1, void* u2, double dt, double nu, double Ri, double Pr,
                    const void* source);
int sdns2d_euler_step(sdns2d_plan* plan, void* u_hat, void* rhs, double dt, double nu, double Ri, double Pr, const void* source);
int sdns2d_ab2_step(sdns2d_plan* plan, void* u_hat, void* u1, void* rhs, double dt, int tstep, double nu, double Ri, double Pr,
                    const void* source);
/* cross2(c, K, u_hat) for a 2-D field: scalar c = 1j*(K0 u1 - K1 u0)  (cython_maths.in:105-147; NS2D.py:21-22) */
int sdns2d_cross2(sdns2d_plan* plan, void* c, const void* u_hat);
/* add_pressure_diffusion_NS2D / add_pressure_diffusion_Bq2D on their own, in place on du (cython_solvers.in:82-127) */
int sdns2d_add_pressure_diffusion(sdns2d_plan* plan, void* du, const void* u_hat, double nu, double Ri, double Pr, void* p_hat);

#ifdef __cplusplus
}
#endif
#endif /* SDNS_B200_H */
