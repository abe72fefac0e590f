"""Slab decomposition bookkeeping (host side), the Python mirror of build_spaces() in
csrc/sdns_api.cu.  One process per GPU; rank r owns spectral k1 in [r*N1l, (r+1)*N1l) and physical
x0 in [r*M0l, (r+1)*M0l) -- the layout of the reference's slab runs (spectralDNS3D_short.py:28-29:
U (3, Np, N, N), U_hat (3, N, Np, N//2+1)).

The kernels never pack: the pass in front of a transpose stores element (field, i, column) straight
into rank i // chunk.  `backward_destinations` / `forward_destinations` describe exactly that map;
tests/test_slab_cpu.py replays it with numpy + torch.distributed(gloo) on two CPU processes.
"""
import numpy as np


def dealias_cutoff(N):
    return int(np.ceil(2./3.*(N//2+1))) - 1


class SlabLayout(object):
    def __init__(self, N, nranks, rank, dealias='2/3-rule'):
        self.N = tuple(int(n) for n in N)
        self.P, self.rank = int(nranks), int(rank)
        N0, N1, N2 = self.N
        self.Nh = N2//2+1
        self.M = tuple(3*n//2 for n in self.N) if dealias == '3/2-rule' else self.N
        if N1 % self.P or self.M[0] % self.P or N0 % self.P:
            raise ValueError('N[0], N[1] (and 3N[0]/2) must be divisible by the number of ranks')
        self.N1l = N1//self.P
        self.M0l = self.M[0]//self.P
        self.k1_slice = slice(rank*self.N1l, (rank+1)*self.N1l)
        self.x0_slice = slice(rank*self.M0l, (rank+1)*self.M0l)
        # kept axis-1 modes entering backward (global memory indices), compact order
        if dealias == '2/3-rule':
            kc = dealias_cutoff(N1)
            if 2*kc+1 < N1:
                kept = np.r_[0:kc+1, N1-kc:N1]
            else:
                kept = np.arange(N1)
            kc2 = dealias_cutoff(N2)
            self.K2n = kc2+1 if kc2+1 < self.Nh else self.Nh
        else:
            kept = np.arange(N1)
            self.K2n = self.Nh
        self.kept1 = kept
        self.K1n = len(kept)
        mine = np.nonzero((kept >= rank*self.N1l) & (kept < (rank+1)*self.N1l))[0]
        self.K1l = len(mine)
        self.c1off = int(mine[0]) if len(mine) else 0            # global compact index of first owned mode
        self.local_kept = kept[mine] - rank*self.N1l              # local memory indices of owned kept modes

    def spectral_shape(self):
        return (self.N[0], self.N1l, self.Nh)

    def physical_shape(self):
        return (self.M0l, self.M[1], self.M[2])

    def backward_destinations(self):
        """B0 output element (x0 global, compact k1 c) -> (rank, local x0, compact k1): array of
        destination ranks per x0."""
        return np.arange(self.M[0])//self.M0l

    def forward_destinations(self):
        """F1 output element with global k1 -> owning rank."""
        return np.arange(self.N[1])//self.N1l
