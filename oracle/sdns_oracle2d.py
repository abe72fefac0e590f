"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + scipy.fft) of the reference's doubly periodic solvers.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product path (spectraldns_b200/)
never does.  Each function cites the reference lines it restates (paths relative to /root/reference):

  solvers/NS2D.py:13-51   get_context / get_curl / get_divergence / Conv (u x curl with a scalar curl)
  solvers/Bq2D.py:13-186  Boussinesq: state (u0, u1, rho), Conv :121-137, add_pressure_diffusion :139-156, ComputeRHS :158-186
  solvers/NS.py:203-261   add_pressure_diffusion and ComputeRHS that NS2D inherits (from .NS import *)
  maths/integrators.py:150-175  RK4 / ForwardEuler / AB2

Pinned by: the reference's known answer for the 2-D Taylor-Green vortex (tests/TG2D.py:12-16, 41-52: kinetic energy
follows exp(-2 nu t)**2 to params.ntol digits; tests/test_NS2D.py drives it for the 2/3-rule, the 3/2-rule and two
meshes) and by the reference's COMPILED 2-D kernels (optimization/cython_solvers.in:82-127
add_pressure_diffusion_Bq2D / _NS2D, cython_maths.in:89-147 cross1_2D / cross2_2D) built into oracle/_ref
(tests/test_oracle.py).  The transform conventions (forward = rfft2 / prod(M), 2/3 truncation on the input of
backward, 3/2 zero padding) are those of the 3-D restatement (oracle/sdns_oracle.py), with the same "parity
unpinned" caveat for the exact 2/3-rule cutoff index, which lives in shenfun.
"""
import numpy as np
import scipy.fft as sfft

from sdns_oracle import dealias_cutoff

__all__ = ['Oracle2D', 'taylor_green_2d']


class Oracle2D(object):
    def __init__(self, N, L=(2*np.pi,)*2, precision='double', dealias='2/3-rule', mask_nyquist=True, kcut=None):
        self.N = tuple(int(n) for n in N)
        self.L = tuple(float(l) for l in L)
        self.float, self.complex = {'single': (np.float32, np.complex64), 'double': (np.float64, np.complex128)}[precision]
        self.dealias = dealias
        N0, N1 = self.N
        self.Nh = N1//2+1
        self.sshape = (N0, self.Nh)
        pf = 1.5 if dealias == '3/2-rule' else 1
        self.M = tuple(int(np.floor(n*pf)) for n in self.N)
        k0 = np.fft.fftfreq(N0, 1./N0)
        k1 = np.fft.rfftfreq(N1, 1./N1)
        self.kint = (k0, k1)
        self.K = [(k0*2*np.pi/self.L[0]).reshape(N0, 1).astype(self.float),
                  (k1*2*np.pi/self.L[1]).reshape(1, self.Nh).astype(self.float)]
        self.K2 = np.zeros(self.sshape, dtype=self.float)
        for i in range(2):
            self.K2 += self.K[i]*self.K[i]
        self.K_over_K2 = np.zeros((2,)+self.sshape, dtype=self.float)
        for i in range(2):
            self.K_over_K2[i] = self.K[i]/np.where(self.K2 == 0, 1, self.K2)
        self.mask = None
        if mask_nyquist:                                   # solvers/NS.py:34
            m = np.ones(self.sshape, dtype=int)
            m[N0//2, :] = 0
            m[:, N1//2] = 0
            self.mask = m
        dm = np.ones(self.sshape, dtype=bool)
        for i, k in enumerate(self.kint):
            s = [1, 1]
            s[i] = len(k)
            kc = dealias_cutoff(self.N[i]) if (kcut is None or kcut[i] is None or kcut[i] < 0) else int(kcut[i])
            dm = dm & (np.abs(k.reshape(s)) <= kc)
        self.dealias_mask = dm

    # ---- transforms
    def forward(self, u, padded=False):
        M = self.M if padded else self.N
        N = self.N
        full = (sfft.rfftn(u, axes=(-2, -1))/np.prod(M)).astype(self.complex)
        if M != N:
            n0, nh = N[0], self.Nh
            h0 = n0//2
            out = np.zeros(u.shape[:-2]+self.sshape, dtype=self.complex)
            out[..., :h0, :] = full[..., :h0, :nh]
            out[..., h0:, :] = full[..., M[0]-(n0-h0):, :nh]
            full = out
        return full

    def backward(self, u_hat, padded=False, dealias=False):
        M = self.M if padded else self.N
        N = self.N
        a = np.asarray(u_hat)
        if M != N:
            n0, nh = N[0], self.Nh
            h0 = n0//2
            full = np.zeros(a.shape[:-2]+(M[0], M[1]//2+1), dtype=a.dtype)
            full[..., :h0, :nh] = a[..., :h0, :]
            full[..., M[0]-(n0-h0):, :nh] = a[..., h0:, :]
            a = full
        elif dealias:
            a = a*self.dealias_mask
        return (sfft.irfftn(a, s=M, axes=(-2, -1))*np.prod(M)).astype(self.float)

    def _bwd_p(self, u_hat):
        if self.dealias == '3/2-rule':
            return self.backward(u_hat, padded=True)
        return self.backward(u_hat, dealias=(self.dealias == '2/3-rule'))

    def _fwd_p(self, u):
        return self.forward(u, padded=(self.dealias == '3/2-rule'))

    # ---- operators
    def cross2(self, u_hat):
        """Scalar curl_hat = 1j*(K0 u1 - K1 u0)  (NS2D.py:20-23; cython_maths.in:105-147)."""
        return (1j*(self.K[0]*u_hat[1] - self.K[1]*u_hat[0])).astype(self.complex)

    def energy_fourier(self, u_hat):
        a = np.asarray(u_hat)
        w = a.real.astype(np.float64)**2 + a.imag.astype(np.float64)**2
        return 2*np.sum(w[..., 1:-1]) + np.sum(w[..., 0]) + np.sum(w[..., -1])

    # ---- NS2D
    def ns2d_conv(self, u_hat):
        """NS2D.py:40-48."""
        curl = self._bwd_p(self.cross2(u_hat))
        u = self._bwd_p(u_hat)
        rhs = np.zeros((2,)+self.sshape, dtype=self.complex)
        rhs[0] = self._fwd_p(u[1]*curl)
        rhs[1] = self._fwd_p(-u[0]*curl)
        return rhs

    def add_pressure_diffusion_ns2d(self, rhs, u_hat, nu):
        """NS.py:203-217 in two dimensions (cython_solvers.in:105-127).  Returns (rhs, P_hat)."""
        nu = self.float(nu)
        P_hat = np.sum(rhs*self.K_over_K2, 0)
        for i in range(2):
            rhs[i] -= P_hat*self.K[i]
        rhs -= nu*self.K2*u_hat
        return rhs, P_hat

    def ns2d_rhs(self, u_hat, nu, source=None, return_p=False):
        """NS.py:219-261 with NS2D's Conv."""
        rhs = self.ns2d_conv(u_hat)
        if self.mask is not None:
            rhs *= self.mask
        rhs, P_hat = self.add_pressure_diffusion_ns2d(rhs, u_hat, nu)
        if source is not None:
            rhs += source
        rhs = rhs.astype(self.complex)
        return (rhs, P_hat.astype(self.complex)) if return_p else rhs

    # ---- Bq2D
    def bq2d_conv(self, ur_hat):
        """Bq2D.py:121-137."""
        ur = self._bwd_p(ur_hat)
        u, rho = ur[:2], ur[2]
        curl = self._bwd_p(self.cross2(ur_hat[:2]))
        rhs = np.zeros((3,)+self.sshape, dtype=self.complex)
        rhs[0] = self._fwd_p(u[1]*curl)
        rhs[1] = self._fwd_p(-u[0]*curl)
        F0 = self._fwd_p(u[0]*rho)
        F1 = self._fwd_p(u[1]*rho)
        rhs[2] = -1j*(self.K[0]*F0 + self.K[1]*F1)
        return rhs

    def add_pressure_diffusion_bq2d(self, rhs, ur_hat, nu, Ri, Pr):
        """Bq2D.py:139-156 (cython_solvers.in:82-103).  Returns (rhs, P_hat)."""
        nu, Ri, Pr = self.float(nu), self.float(Ri), self.float(Pr)
        u_hat, rho_hat = ur_hat[:2], ur_hat[2]
        P_hat = np.sum(rhs[:2]*self.K_over_K2, 0)
        P_hat -= Ri*rho_hat*self.K_over_K2[1]
        for i in range(2):
            rhs[i] -= P_hat*self.K[i]
        rhs[0] -= nu*self.K2*u_hat[0]
        rhs[1] -= (nu*self.K2*u_hat[1] + Ri*rho_hat)
        rhs[2] -= nu*self.K2*rho_hat/Pr
        return rhs, P_hat

    def bq2d_rhs(self, ur_hat, nu, Ri, Pr, return_p=False):
        """Bq2D.py:158-186."""
        rhs = self.bq2d_conv(ur_hat)
        if self.mask is not None:
            rhs *= self.mask
        rhs, P_hat = self.add_pressure_diffusion_bq2d(rhs, ur_hat, nu, Ri, Pr)
        rhs = rhs.astype(self.complex)
        return (rhs, P_hat.astype(self.complex)) if return_p else rhs

    # ---- integrators (maths/integrators.py:150-175)
    def rk4_step(self, u0, rhs_fn, dt):
        a = np.array([1./6., 1./3., 1./3., 1./6.], dtype=self.float)
        b = np.array([0.5, 0.5, 1.], dtype=self.float)
        dt = self.float(dt)
        u0 = u0.copy()
        u1 = u0.copy()
        u2 = u0.copy()
        for rk in range(4):
            rhs = rhs_fn(u0)
            if rk < 3:
                u0 = (u1 + b[rk]*dt*rhs).astype(self.complex)
            u2 = (u2 + a[rk]*dt*rhs).astype(self.complex)
        return u2

    def forward_euler_step(self, u0, rhs_fn, dt):
        return (u0 + rhs_fn(u0)*self.float(dt)).astype(self.complex)

    def ab2_step(self, u0, u1, rhs_fn, dt, tstep):
        rhs = rhs_fn(u0)*self.float(dt)
        u0 = (u0 + rhs) if tstep == 0 else (u0 + (1.5*rhs - 0.5*u1))
        return u0.astype(self.complex), rhs.astype(self.complex)

    def solve(self, u_hat, solver, nsteps, dt, nu, Ri=0.1, Pr=1.0):
        fn = (lambda v: self.ns2d_rhs(v, nu)) if solver == 'NS2D' else (lambda v: self.bq2d_rhs(v, nu, Ri, Pr))
        u = np.asarray(u_hat).astype(self.complex)
        for _ in range(nsteps):
            u = self.rk4_step(u, fn, dt)
        return u

    def mesh(self, padded=False):
        M = self.M if padded else self.N
        return [np.arange(M[0], dtype=float).reshape(M[0], 1)*self.L[0]/M[0],
                np.arange(M[1], dtype=float).reshape(1, M[1])*self.L[1]/M[1]]


def taylor_green_2d(o):
    """tests/TG2D.py:12-16."""
    X = o.mesh()
    U = np.zeros((2,)+o.N, dtype=o.float)
    U[0] = np.sin(X[0])*np.cos(X[1])
    U[1] = -np.sin(X[1])*np.cos(X[0])
    return o.forward(U)
