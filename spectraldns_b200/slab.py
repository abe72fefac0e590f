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
            raise ValueError('this numpy replay covers the even split only: N[0], N[1] (and 3N[0]/2) divisible by the number of ranks')
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


# ---------------------------------------------------------------------------------------------
# Copy-engine exchange (csrc/sdns_api.cu: chunk_bound, k1_chunk, b0_chunk_ce, rhs_ce): the passes in front of a
# transpose run in chunks; a chunk writes the part of every peer into a send slot and one strided 2-D copy per
# peer moves it.  The functions below are the host mirror of that index arithmetic (elements, not bytes);
# tests/test_slab_cpu.py replays them with numpy buffers and gloo messages.
# ---------------------------------------------------------------------------------------------
def chunk_bound(n, nc, c):
    """Boundary c (0..nc) of an axis of n entries cut into nc chunks with weights 3,..,3,2,1."""
    w = [min(nc - i, 3) for i in range(nc)]
    return (n*sum(w[:c]) + sum(w)//2)//sum(w)


class Copy2D(object):
    """One cudaMemcpy2DAsync: `height` rows of `width` elements, row r at src_off + r*spitch -> dst_off + r*dpitch."""
    def __init__(self, src_off, spitch, dst_off, dpitch, width, height):
        self.src_off, self.spitch, self.dst_off, self.dpitch = int(src_off), int(spitch), int(dst_off), int(dpitch)
        self.width, self.height = int(width), int(height)

    def gather(self, src_flat):
        rows = self.src_off + self.spitch*np.arange(self.height)[:, None] + np.arange(self.width)[None, :]
        return src_flat[rows]

    def scatter(self, dst_flat, payload):
        rows = self.dst_off + self.dpitch*np.arange(self.height)[:, None] + np.arange(self.width)[None, :]
        dst_flat[rows] = payload


class ExchangePlan(object):
    """Send-slot layouts and per-chunk copies of rank `L.rank` (L: SlabLayout) for nf fields."""
    def __init__(self, L, nchunk=4):
        self.L, self.nc = L, int(nchunk)
        self.K2p = (L.K2n + 1) & ~1
        self.Nhp = (L.Nh + 1) & ~1

    # B0: slot r holds (6|nf, M0l, K1l, K2p); destination array W0 is (nf, M0l, K1n, K2p) on rank r
    def b0_slot(self, nf):
        return nf*self.L.M0l*self.L.K1l*self.K2p

    def b0_kept(self, c):
        return chunk_bound(self.L.K1l, self.nc, c), chunk_bound(self.L.K1l, self.nc, c+1)

    def b0_send_index(self, nf, r, f, x0l, c1, c2):
        """element (field f, local x0 of rank r, this rank's compact mode c1, k2) inside the send buffer"""
        L = self.L
        return r*self.b0_slot(nf) + ((f*L.M0l + x0l)*L.K1l + c1)*self.K2p + c2

    def b0_copy(self, nf, r, c):
        L = self.L
        ka, kb = self.b0_kept(c)
        return Copy2D(r*self.b0_slot(nf) + ka*self.K2p, L.K1l*self.K2p,
                      (L.c1off + ka)*self.K2p, L.K1n*self.K2p, (kb - ka)*self.K2p, nf*L.M0l)

    # F1: slot r holds (nf, N1l, M0l, Nhp); destination array W3 is (nf, N1l, M0, Nhp) on rank r
    def f1_slot(self, nf):
        L = self.L
        return nf*L.N1l*L.M0l*self.Nhp

    def f1_planes(self, c):
        return chunk_bound(self.L.M0l, self.nc, c), chunk_bound(self.L.M0l, self.nc, c+1)

    def f1_send_index(self, nf, r, f, k1l, x0l, c2):
        L = self.L
        return r*self.f1_slot(nf) + ((f*L.N1l + k1l)*L.M0l + x0l)*self.Nhp + c2

    def f1_copy(self, nf, r, c):
        L = self.L
        xa, xb = self.f1_planes(c)
        return Copy2D(r*self.f1_slot(nf) + xa*self.Nhp, L.M0l*self.Nhp,
                      (L.rank*L.M0l + xa)*self.Nhp, L.M[0]*self.Nhp, (xb - xa)*self.Nhp, nf*L.N1l)
