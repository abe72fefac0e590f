"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the triply periodic pseudo-spectral hot path.

A dependency-free (numpy + scipy.fft) restatement of what the reference computes on the path
BASELINE.json names: the NS / VV / MHD right-hand side and the RK4 step.  Every function cites
the reference file:line it follows (paths relative to /root/reference).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
file; the product package (spectraldns_b200/) never does.

Parity pinning: this restatement is checked (tests/test_oracle.py, oracle/make_golden.py)
  (1) against the reference's own known answers: TG k=0.124953117517, w=0.375249930801
      (tests/TG.py:125-126), TG-MHD k=0.124565408177, b=0.124637762143 (tests/TGMHD.py:25-26);
  (2) field-level against the reference's UNMODIFIED solvers/{NS,VV,MHD}.py +
      maths/integrators.py imported from /root/reference over the numpy stand-in substrate in
      oracle/shim (fixtures tests/golden/*.npz written by oracle/make_golden.py);
  (3) against the reference's own compiled kernels for this path (optimization/cython_{maths,solvers,
      integrators}.in built into oracle/_ref by oracle/build_ref_cython.py): cross1, cross2,
      add_pressure_diffusion_NS, RK4, ForwardEuler, AB2 (tests/test_oracle.py).
The FFT / padding / truncation arithmetic itself lives in shenfun + mpi4py-fft (pins
shenfun>=4.0.2, mpi4py-fft>=2.0.3, setup.py:86, conf/conda/meta.yaml:34-35), which are NOT in
/root/reference; their published algorithm is restated here and the exact 2/3-rule cutoff index
and Nyquist weighting of padded transforms are therefore "parity unpinned" (SURVEY.md section 8c):
the in-tree convention |k_i| < 2/3*(N_i/2+1) of spectralDNS3D_short.py:44-46 is used.
"""
import numpy as np
import scipy.fft as sfft

__all__ = ['Oracle', 'dealias_cutoff', 'taylor_green', 'taylor_green_mhd', 'isotropic_field']


def dealias_cutoff(N):
    """Largest kept integer |k| under |k| < 2/3*(N/2+1) (spectralDNS3D_short.py:44-46)."""
    return int(np.ceil(2./3.*(N//2+1))) - 1


class Oracle(object):
    """State-free helper bound to one grid/precision/dealias configuration.

    Layouts follow solvers/NS.py:51-63: physical (ncomp, M0, M1, M2) real, spectral
    (ncomp, N0, N1, N2//2+1) complex, C order.
    """

    def __init__(self, N, L=(2*np.pi,)*3, precision='double', dealias='2/3-rule',
                 mask_nyquist=True, workers=-1, kcut=None):
        self.N = tuple(int(n) for n in N)
        self.L = tuple(float(l) for l in L)
        # solvers/spectralinit.py:25-29
        self.float, self.complex = {'single': (np.float32, np.complex64),
                                    'double': (np.float64, np.complex128)}[precision]
        self.dealias = dealias
        self.workers = workers
        N0, N1, N2 = self.N
        self.Nh = N2//2+1
        self.sshape = (N0, N1, self.Nh)
        # solvers/NS.py:29-31: padded physical shape for the 3/2-rule
        pf = 1.5 if dealias == '3/2-rule' else 1
        self.M = tuple(int(np.floor(n*pf)) for n in self.N)
        # solvers/NS.py:38-48 (wavenumbers scaled by 2*pi/L; Nyquist of the r2c axis is +N/2)
        k0 = np.fft.fftfreq(N0, 1./N0)
        k1 = np.fft.fftfreq(N1, 1./N1)
        k2 = np.fft.rfftfreq(N2, 1./N2)
        self.kint = (k0, k1, k2)
        self.K = [(k0*2*np.pi/self.L[0]).reshape(N0, 1, 1).astype(self.float),
                  (k1*2*np.pi/self.L[1]).reshape(1, N1, 1).astype(self.float),
                  (k2*2*np.pi/self.L[2]).reshape(1, 1, self.Nh).astype(self.float)]
        self.K2 = np.zeros(self.sshape, dtype=self.float)
        for i in range(3):
            self.K2 += self.K[i]*self.K[i]
        self.K_over_K2 = np.zeros((3,)+self.sshape, dtype=self.float)
        for i in range(3):
            self.K_over_K2[i] = self.K[i]/np.where(self.K2 == 0, 1, self.K2)
        # solvers/NS.py:34 -- zero every mode with an index N_i/2
        self.mask = None
        if mask_nyquist:
            m = np.ones(self.sshape, dtype=int)
            for i, n in enumerate(self.N):
                if n % 2 == 0:
                    s = [slice(None)]*3
                    s[i] = n//2
                    m[tuple(s)] = 0
            self.mask = m
        # 2/3-rule truncation mask applied to the input of backward (solvers/NS.py:29-31,
        # cutoff spectralDNS3D_short.py:44-46)
        dm = np.ones(self.sshape, dtype=bool)
        for i, k in enumerate(self.kint):
            s = [1, 1, 1]
            s[i] = len(k)
            kc = dealias_cutoff(self.N[i]) if (kcut is None or kcut[i] is None or kcut[i] < 0) else int(kcut[i])
            dm = dm & (np.abs(k.reshape(s)) <= kc)        # kcut: the plan parameter sdns_config.kcut (tests only)
        self.dealias_mask = dm

    # ---- transforms (shenfun TensorProductSpace.forward/backward as used at NS.py:93,103,128,135)
    def forward(self, u, padded=False):
        """rfftn(u)/prod(M); with 3/2-rule padding keep the N-sized corner blocks."""
        M = self.M if padded else self.N
        N = self.N
        full = sfft.rfftn(u, axes=(-3, -2, -1), workers=self.workers)
        full = (full/np.prod(M)).astype(self.complex)
        if M != N:
            n0, n1, nh = N[0], N[1], self.Nh
            h0, h1 = n0//2, n1//2
            out = np.zeros(u.shape[:-3]+self.sshape, dtype=self.complex)
            out[..., :h0, :h1, :] = full[..., :h0, :h1, :nh]
            out[..., :h0, h1:, :] = full[..., :h0, M[1]-(n1-h1):, :nh]
            out[..., h0:, :h1, :] = full[..., M[0]-(n0-h0):, :h1, :nh]
            out[..., h0:, h1:, :] = full[..., M[0]-(n0-h0):, M[1]-(n1-h1):, :nh]
            full = out
        return full

    def backward(self, u_hat, padded=False, dealias=False):
        """irfftn*prod(M). padded: scatter corner blocks into the 3/2 spectrum; dealias: apply the
        2/3 truncation to the input (shenfun dealias_direct)."""
        M = self.M if padded else self.N
        N = self.N
        a = np.asarray(u_hat)
        if M != N:
            n0, n1, nh = N[0], N[1], self.Nh
            h0, h1 = n0//2, n1//2
            full = np.zeros(a.shape[:-3]+(M[0], M[1], M[2]//2+1), dtype=a.dtype)
            full[..., :h0, :h1, :nh] = a[..., :h0, :h1, :]
            full[..., :h0, M[1]-(n1-h1):, :nh] = a[..., :h0, h1:, :]
            full[..., M[0]-(n0-h0):, :h1, :nh] = a[..., h0:, :h1, :]
            full[..., M[0]-(n0-h0):, M[1]-(n1-h1):, :nh] = a[..., h0:, h1:, :]
            a = full
        elif dealias:
            a = a*self.dealias_mask
        r = sfft.irfftn(a, s=M, axes=(-3, -2, -1), workers=self.workers)*np.prod(M)
        return r.astype(self.float)

    def _bwd_p(self, u_hat):
        """Backward on the dealiased space Tp (solvers/NS.py:29-32)."""
        if self.dealias == '3/2-rule':
            return self.backward(u_hat, padded=True)
        return self.backward(u_hat, dealias=(self.dealias == '2/3-rule'))

    def _fwd_p(self, u):
        return self.forward(u, padded=(self.dealias == '3/2-rule'))

    # ---- pointwise pieces
    def cross1(self, a, b):
        """maths/cross.py:16-28; optimization/cython_maths.in:13-30"""
        c = np.empty_like(a)
        c[0] = a[1]*b[2] - a[2]*b[1]
        c[1] = a[2]*b[0] - a[0]*b[2]
        c[2] = a[0]*b[1] - a[1]*b[0]
        return c

    def cross2(self, a, b):
        """c = 1j*(a x b), a real (list of broadcast K or dense), b complex
        (maths/cross.py:30-35; cython_maths.in:32-86)"""
        c = np.empty_like(b)
        c[0] = a[1]*b[2] - a[2]*b[1]
        c[1] = a[2]*b[0] - a[0]*b[2]
        c[2] = a[0]*b[1] - a[1]*b[0]
        c *= 1j
        return c

    def energy_fourier(self, u_hat):
        """shenfun.fourier.energy_fourier as used at tests/TG.py:101: Hermitian-weighted sum |u_hat|^2."""
        a = np.asarray(u_hat)
        w = a.real.astype(np.float64)**2 + a.imag.astype(np.float64)**2
        if self.N[2] % 2 == 0:
            return 2*np.sum(w[..., 1:-1]) + np.sum(w[..., 0]) + np.sum(w[..., -1])
        return 2*np.sum(w[..., 1:]) + np.sum(w[..., 0])

    # ---- diagnostics and forcing of demo/Isotropic.py (user code of the reference, restated for the device versions)
    def spectrum(self, u_hat):
        """demo/Isotropic.py:88-118 on the global array: (Ek, bins).  uiui counts the first and last k2 plane once
        and the others twice, times 4 pi / 3; shell i holds np.digitize(sqrt(K2), bins, right=True) == i + 1."""
        u_hat = np.asarray(u_hat)
        uiui = np.zeros(u_hat[0].shape)
        uiui[..., 1:-1] = 2*np.sum((u_hat[..., 1:-1]*np.conj(u_hat[..., 1:-1])).real, axis=0)
        uiui[..., 0] = np.sum((u_hat[..., 0]*np.conj(u_hat[..., 0])).real, axis=0)
        uiui[..., -1] = np.sum((u_hat[..., -1]*np.conj(u_hat[..., -1])).real, axis=0)
        uiui *= (4./3.*np.pi)
        Nb = int(np.sqrt(sum((np.array(self.N)/2)**2)/3))
        bins = np.array(range(0, Nb))+0.5
        z = np.digitize(np.sqrt(self.K2), bins, right=True)
        Ek = np.zeros(Nb)
        ll = np.zeros(Nb)
        for i, k in enumerate(bins[1:]):
            k0 = bins[i]
            ii = np.where((z > k0) & (z <= k))
            ll[i] = len(ii[0])
            Ek[i] = (k**3 - k0**3)*np.sum(uiui[ii])
        for i in range(Nb):
            if not ll[i] == 0:
                Ek[i] = Ek[i] / ll[i]
        return Ek, bins

    def forcing_rescale(self, u_hat, Kf2, target_energy):
        """The low-wavenumber forcing of demo/Isotropic.py:161-184 on the global array: returns (u_hat rescaled in
        place, energy_new, energy_lower, alpha)."""
        k2_mask = np.where(self.K2 <= Kf2**2, 1, 0)
        u_hat[:, 0, 0, 0] = 0
        energy_new = self.energy_fourier(u_hat)
        energy_lower = self.energy_fourier(u_hat*k2_mask)
        energy_upper = energy_new - energy_lower
        alpha = np.sqrt((target_energy - energy_upper)/energy_lower)
        u_hat *= (alpha*k2_mask + (1-k2_mask))
        return u_hat, self.energy_fourier(u_hat), energy_lower, alpha

    def enstrophy(self, u_hat):
        """dissipation = energy_fourier(cross2(K, U_hat)) (demo/Isotropic.py:243-244)."""
        return self.energy_fourier(self.cross2(self.K, u_hat))

    def divergence_norm(self, u_hat):
        """L2_norm(get_divergence(...)) (demo/Isotropic.py:78-86, 245-247; solvers/NS.py:107-110): mean of div(u)^2."""
        d = self.backward(1j*(self.K[0]*u_hat[0] + self.K[1]*u_hat[1] + self.K[2]*u_hat[2]))
        return float(np.sum(d.astype(np.float64)**2)/np.prod(self.N))

    # ---- NS (solvers/NS.py)
    def ns_conv(self, u_hat, convection='Vortex'):
        K = self.K
        if convection == 'Vortex':
            # NS.py:191-198: u x curl(u) on the dealiased space
            u = self._bwd_p(u_hat)
            curl = self._bwd_p(self.cross2(K, u_hat))         # NS.py:124-129
            return self._fwd_p(self.cross1(u, curl))          # NS.py:131-136
        u = self._bwd_p(u_hat)
        rhs = np.zeros_like(u_hat)
        if convection in ('Standard', 'Skewed'):
            # NS.py:138-145
            for i in range(3):
                g = np.stack([self._bwd_p(1j*K[j]*u_hat[i]) for j in range(3)])
                rhs[i] = self._fwd_p(np.sum(u*g, 0))
        if convection in ('Divergence', 'Skewed'):
            # NS.py:147-162
            UU = np.stack([self._fwd_p(u[0]*u[i]) for i in range(3)])
            rhs[0] += 1j*(K[0]*UU[0] + K[1]*UU[1] + K[2]*UU[2])
            rhs[1] += 1j*K[0]*UU[1]
            rhs[2] += 1j*K[0]*UU[2]
            UU = np.stack([self._fwd_p(u[1]*u[1]), self._fwd_p(u[1]*u[2]), self._fwd_p(u[2]*u[2])])
            rhs[1] += 1j*K[1]*UU[0] + 1j*K[2]*UU[1]
            rhs[2] += 1j*K[1]*UU[1] + 1j*K[2]*UU[2]
        rhs *= (-0.5 if convection == 'Skewed' else -1)       # NS.py:170,178,188
        return rhs

    def add_pressure_diffusion(self, rhs, u_hat, nu):
        """solvers/NS.py:203-217 (Cython: add_pressure_diffusion_NS_, cython_solvers.in:44-80):
        P_hat = sum_i rhs_i K_i/K^2 ; rhs_i -= P_hat K_i + nu K^2 u_hat_i.  Returns (rhs, P_hat); rhs is
        updated in place."""
        nu = self.float(nu)
        P_hat = np.sum(rhs*self.K_over_K2, 0)
        for i in range(3):
            rhs[i] -= P_hat*self.K[i]
        rhs -= nu*self.K2*u_hat
        return rhs, P_hat

    def ns_rhs(self, u_hat, nu, convection='Vortex', source=None, return_p=False):
        """solvers/NS.py:219-261 with add_pressure_diffusion NS.py:203-217
        (cython_solvers.in:44-80)."""
        nu = self.float(nu)
        rhs = self.ns_conv(u_hat, convection)
        if self.mask is not None:
            rhs *= self.mask                                   # NS.py:253-254
        rhs, P_hat = self.add_pressure_diffusion(rhs, u_hat, nu)
        if source is not None:
            rhs += source                                      # NS.py:259
        rhs = rhs.astype(self.complex)
        if return_p:
            return rhs, P_hat.astype(self.complex)
        return rhs

    # ---- VV (solvers/VV.py)
    def vv_rhs(self, w_hat, nu, source=None):
        """solvers/VV.py:92-100 (Conv), :52-67 (compute_velocity), :105-110 (add_linear), :112-146."""
        nu = self.float(nu)
        v_hat = self.cross2(self.K_over_K2, w_hat)            # u_hat = i k x w_hat / k^2
        u = self._bwd_p(v_hat)
        w = self._bwd_p(w_hat)
        v_hat = self._fwd_p(self.cross1(u, w))
        rhs = self.cross2(self.K, v_hat)
        if self.mask is not None:
            rhs *= self.mask
        rhs -= nu*self.K2*w_hat
        if source is not None:
            rhs += source
        return rhs.astype(self.complex)

    # ---- MHD (solvers/MHD.py)
    def mhd_rhs(self, ub_hat, nu, eta):
        """solvers/MHD.py:119-127 (Conv), :99-110, :89-97 (set_Elsasser), :132-149, :151-176."""
        nu, eta = self.float(nu), self.float(eta)
        K = self.K
        ub = self._bwd_p(ub_hat)
        z0 = ub[:3] + ub[3:]
        z1 = ub[:3] - ub[3:]
        ZZ = np.empty((3, 3)+self.sshape, dtype=self.complex)
        for i in range(3):
            for j in range(3):
                ZZ[i, j] = self._fwd_p(z0[i]*z1[j])
        rhs = np.empty_like(ub_hat)
        rhs[:3] = -1j*(K[0]*(ZZ[:, 0] + ZZ[0, :]) + K[1]*(ZZ[:, 1] + ZZ[1, :])
                       + K[2]*(ZZ[:, 2] + ZZ[2, :]))/2.0
        rhs[3:] = 1j*(K[0]*(ZZ[0, :] - ZZ[:, 0]) + K[1]*(ZZ[1, :] - ZZ[:, 1])
                      + K[2]*(ZZ[2, :] - ZZ[:, 2]))/2.0
        if self.mask is not None:
            rhs *= self.mask
        P_hat = np.sum(rhs[:3]*self.K_over_K2, 0)
        for i in range(3):
            rhs[i] -= P_hat*K[i]
        rhs[:3] -= nu*self.K2*ub_hat[:3]
        rhs[3:] -= eta*self.K2*ub_hat[3:]
        return rhs.astype(self.complex)

    # ---- integrators (maths/integrators.py)
    def rk4_step(self, u0, rhs_fn, dt):
        """maths/integrators.py:150-159 with a,b of :185-186 cast to context.float."""
        a = np.array([1./6., 1./3., 1./3., 1./6.], dtype=self.float)
        b = np.array([0.5, 0.5, 1.], dtype=self.float)
        dt = self.float(dt)
        u0 = u0.copy()
        u1 = u0.copy()
        u2 = u0.copy()
        for rk in range(4):
            rhs = rhs_fn(u0)
            if rk < 3:
                u0[:] = u1 + b[rk]*dt*rhs
            u2 += a[rk]*dt*rhs
        u0[:] = u2
        return u0

    def forward_euler_step(self, u0, rhs_fn, dt):
        """maths/integrators.py:161-165"""
        return (u0 + rhs_fn(u0)*self.float(dt)).astype(self.complex)

    def ab2_step(self, u0, u1, rhs_fn, dt, tstep):
        """maths/integrators.py:167-175; returns (u0_new, u1_new)."""
        dt = self.float(dt)
        rhs = rhs_fn(u0)
        if tstep == 0:
            un = u0 + rhs*dt
        else:
            un = u0 + (1.5*rhs*dt - 0.5*u1)
        return un.astype(self.complex), (rhs*dt).astype(self.complex)

    def bs5_solve(self, u0, rhs_fn, dt, T, adaptive, TOL=1e-6):
        """maths/integrators.py:15-147 (adaptiveRK) with the BS5 tableau of :199-208 driven by the
        time loop of spectralDNS/__init__.py:94-111 and NS.end_of_tstep (solvers/NS.py:112-122).
        Returns (u, number of steps, t)."""
        A = np.array([[0, 0, 0, 0, 0, 0, 0, 0],
                      [1/6, 0, 0, 0, 0, 0, 0, 0],
                      [2/27, 4/27, 0, 0, 0, 0, 0, 0],
                      [183/1372, -162/343, 1053/1372, 0, 0, 0, 0, 0],
                      [68/297, -4/11, 42/143, 1960/3861, 0, 0, 0, 0],
                      [597/22528, 81/352, 63099/585728, 58653/366080, 4617/20480, 0, 0, 0],
                      [174197/959244, -30942/79937, 8152137/19744439, 666106/1039181, -29421/29068, 482048/414219, 0, 0],
                      [587/8064, 0, 4440339/15491840, 24353/124800, 387/44800, 2152/5985, 7267/94080, 0]],
                     dtype=self.float)
        b = np.array([587/8064, 0, 4440339/15491840, 24353/124800, 387/44800, 2152/5985, 7267/94080, 0], dtype=self.float)
        bhat = np.array([2479/34992, 0, 123/416, 612941/3411720, 43/1440, 2272/6561, 79937/1113912, 3293/556956],
                        dtype=self.float)
        s = 8
        u0 = np.array(u0, dtype=self.complex)
        fY = np.zeros((s,)+u0.shape, dtype=u0.dtype)
        offset, t, tstep = 0, 0.0, 0
        dt = self.float(dt)
        while t + dt <= T + 1e-12:
            facmax, fac, facmin = 2, 0.8, 0.01
            while True:
                dt_prev = dt
                offset = (offset - 1) % s
                for i in range(s):
                    if tstep == 0 or i != 0:
                        fY[(i+offset) % s] = u0
                        for j in range(i):
                            fY[(i+offset) % s] += dt*A[i, j]*fY[(j+offset) % s]
                        fY[(i+offset) % s] = rhs_fn(fY[(i+offset) % s])
                u_new = u0.copy()
                err = np.zeros_like(u0)
                for j in range(s):
                    u_new += dt*b[j]*fY[(j+offset) % s]
                    err += dt*(b[j]-bhat[j])*fY[(j+offset) % s]
                sc = TOL + np.maximum(np.abs(u0), np.abs(u_new))*TOL
                nsq = np.array([np.sum(np.power(np.abs(err[k]/sc[k]), 2)) for k in range(u0.shape[0])])
                est = np.max(np.sqrt(nsq))/np.sqrt(np.prod(self.sshape))
                factor = min(facmax, max(facmin, fac*pow((1/est), 1.0/5)))
                if adaptive:
                    dt = dt*factor
                    if est > 1.0:
                        facmax = 1
                        offset += 1
                        continue
                break
            u0 = u_new
            t += dt_prev
            tstep += 1
            if abs(t - T) < 1e-12:                      # NS.end_of_tstep
                break
            if abs(t + dt - T) < 1e-12 or t + dt >= T + 1e-12:
                dt = self.float(T - t)
        return u0, tstep, t

    def solve(self, u_hat, solver, nsteps, dt, nu, eta=None, convection='Vortex', source=None):
        """nsteps of the while-loop body of spectralDNS/__init__.py:94-98 with RK4."""
        if solver == 'NS':
            fn = lambda u: self.ns_rhs(u, nu, convection, source)
        elif solver == 'VV':
            fn = lambda u: self.vv_rhs(u, nu, source)
        elif solver == 'MHD':
            fn = lambda u: self.mhd_rhs(u, nu, eta)
        else:
            raise ValueError(solver)
        u = np.array(u_hat, dtype=self.complex)
        for _ in range(nsteps):
            u = self.rk4_step(u, fn, dt)
        return u

    def mesh(self, padded=False):
        """T.local_mesh(True) (solvers/NS.py:37)."""
        M = self.M if padded else self.N
        X = []
        for i in range(3):
            s = [1, 1, 1]
            s[i] = M[i]
            X.append((np.arange(M[i], dtype=float)*self.L[i]/M[i]).reshape(s).astype(self.float))
        return X


# ---- synthetic initial fields (BASELINE.json configs)
def taylor_green(o):
    """tests/TG.py:23-28"""
    X = o.mesh()
    U = np.zeros((3,)+o.N, dtype=o.float)
    U[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
    U[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
    return U


def taylor_green_mhd(o):
    """tests/TGMHD.py:4-12"""
    X = o.mesh()
    UB = np.zeros((6,)+o.N, dtype=o.float)
    UB[0] = np.sin(X[0])*np.cos(X[1])*np.cos(X[2])
    UB[1] = -np.cos(X[0])*np.sin(X[1])*np.cos(X[2])
    UB[3] = np.sin(X[0])*np.sin(X[1])*np.cos(X[2])
    UB[4] = np.cos(X[0])*np.cos(X[1])*np.cos(X[2])
    return UB


def isotropic_field(o, seed=0, Kf2=3, kd=50., Re_lam=84., ncomp=3):
    """Seeded broadband solenoidal field following demo/Isotropic.py:29-76 (Rogallo phases,
    k^-5/3 tail, Nyquist-masked, projected, k=0 mode zeroed, rescaled to the target energy).
    Generated once globally with seed `seed` (the reference seeds per rank, :34)."""
    rng = np.random.RandomState(seed)
    K, K2 = [k.astype(np.float64) for k in o.K], o.K2.astype(np.float64)
    nu = 1./kd**(4./3.)
    out = []
    for _ in range(ncomp//3):
        k2_mask = np.where(K2 <= Kf2**2, 1, 0)
        k = np.sqrt(K2)
        k = np.where(k == 0, 1, k)
        kk = np.where(K2 == 0, 1, K2)
        k1, k2, k3 = K
        ksq = np.sqrt(k1**2+k2**2)
        ksq = np.where(ksq == 0, 1, ksq)
        E0 = np.sqrt(9./11./Kf2*K2/Kf2**2)*k2_mask
        E1 = np.sqrt(9./11./Kf2*(k/Kf2)**(-5./3.))*(1-k2_mask)
        Ek = E0 + E1
        theta1, theta2, phi = rng.random_sample((3,)+o.sshape)*2j*np.pi
        alpha = np.sqrt(Ek/4./np.pi/kk)*np.exp(1j*theta1)*np.cos(phi)
        beta = np.sqrt(Ek/4./np.pi/kk)*np.exp(1j*theta2)*np.sin(phi)
        U_hat = np.zeros((3,)+o.sshape, dtype=np.complex128)
        U_hat[0] = (alpha*k*k2 + beta*k1*k3)/(k*ksq)
        U_hat[1] = (beta*k2*k3 - alpha*k*k1)/(k*ksq)
        U_hat[2] = beta*ksq/k
        if o.mask is not None:
            U_hat *= o.mask
        # make Hermitian-consistent: round trip through physical space (Isotropic.py:56-57)
        od = Oracle(o.N, o.L, 'double', o.dealias, o.mask is not None, o.workers)
        U_hat = od.forward(od.backward(U_hat))
        U_hat -= (K[0]*U_hat[0]+K[1]*U_hat[1]+K[2]*U_hat[2])*od.K_over_K2
        U_hat[:, 0, 0, 0] = 0
        energy = 0.5*od.energy_fourier(U_hat)
        target = Re_lam*(nu*kd)**2/np.sqrt(20./3.)
        U_hat *= np.sqrt(target/energy)
        out.append(U_hat)
    return np.concatenate(out).astype(o.complex)
