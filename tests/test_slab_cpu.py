"""CPU test of the N>1 path's host logic: two gloo processes replay the slab pipeline (axis-0 pass,
peer 'stores' = all_to_all, axis-1 pass, axis-2 pass) in numpy using spectraldns_b200.slab and must
reproduce the single-process oracle transform on their slabs, for the 2/3-rule (pruned, compact
axis-1 set), the 3/2-rule and no dealiasing."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]


def _axis_pad(a, axis, M, N):
    """scatter N spectral entries (fftfreq order) into a length-M axis (corner copy)."""
    if M == N:
        return a
    sh = list(a.shape)
    sh[axis] = M
    out = np.zeros(sh, dtype=a.dtype)
    lo = [slice(None)]*a.ndim
    hi_s = [slice(None)]*a.ndim
    hi_d = [slice(None)]*a.ndim
    lo[axis] = slice(0, N//2)
    hi_s[axis] = slice(N//2, N)
    hi_d[axis] = slice(M-(N-N//2), M)
    out[tuple(lo)] = a[tuple(lo)]
    out[tuple(hi_d)] = a[tuple(hi_s)]
    return out


def _a2a(dist, recv, send, rank, world):
    """all_to_all from point-to-point messages (gloo has no alltoall): what the GPU path does with
    direct peer stores."""
    reqs = []
    for r in range(world):
        if r == rank:
            recv[r].copy_(send[r])
        else:
            reqs.append(dist.isend(send[r], r))
            reqs.append(dist.irecv(recv[r], r))
    for q in reqs:
        q.wait()


def _worker(rank, world, port, dealias, N, ret):
    import torch
    import torch.distributed as dist
    import sdns_oracle as so
    from spectraldns_b200.slab import SlabLayout
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        o = so.Oracle(N, dealias=dealias)
        L = SlabLayout(N, world, rank, dealias)
        rng = np.random.RandomState(4)
        u_hat = o.forward(rng.standard_normal(tuple(N)))               # global spectrum (same on all ranks)
        mine = u_hat[:, L.k1_slice, :]                                  # this rank's slab
        M = L.M
        # ---- B0: axis-0 inverse on owned kept columns, truncation/padding at load
        cols = mine[:, L.local_kept, :L.K2n]
        if dealias == '2/3-rule':
            kc0 = so.dealias_cutoff(N[0])
            m0 = np.abs(np.fft.fftfreq(N[0], 1./N[0])) <= kc0
            cols = cols*m0[:, None, None]
        w = np.fft.ifft(_axis_pad(cols, 0, M[0], N[0]), axis=0)*M[0]    # (M0, K1l, K2n)
        # ---- peer stores: x0 chunk d goes to rank d at compact k1 offset c1off
        dest = L.backward_destinations()
        send = [torch.from_numpy(np.ascontiguousarray(w[dest == d])) for d in range(world)]
        all_layouts = [SlabLayout(N, world, r, dealias) for r in range(world)]
        recv = [torch.zeros((L.M0l, all_layouts[r].K1l, L.K2n), dtype=torch.complex128) for r in range(world)]
        _a2a(dist, recv, send, rank, world)
        W0 = np.zeros((L.M0l, L.K1n, L.K2n), dtype=complex)
        for r in range(world):
            c = all_layouts[r].c1off
            W0[:, c:c+all_layouts[r].K1l] = recv[r].numpy()
        # ---- B1: compact axis 1 -> full, inverse; Z: c2r
        full1 = np.zeros((L.M0l, M[1], L.K2n), dtype=complex)
        if dealias == '3/2-rule':
            full1 = _axis_pad(W0, 1, M[1], N[1])
        else:
            full1[:, L.kept1] = W0
        w1 = np.fft.ifft(full1, axis=1)*M[1]
        zin = np.zeros((L.M0l, M[1], M[2]//2+1), dtype=complex)
        zin[..., :L.K2n] = w1
        phys = np.fft.irfft(zin, n=M[2], axis=2)*M[2]
        ref = o._bwd_p(u_hat)[L.x0_slice]
        e_b = np.linalg.norm(phys-ref)/np.linalg.norm(ref)
        # ---- forward: Z r2c, F1 axis 1 (+truncation), stores to the k1 owner, F0 axis 0
        v = rng.standard_normal(o.M)                                    # global physical field
        z = np.fft.rfft(v[L.x0_slice], axis=2)[..., :L.Nh]/np.prod(M)
        f1 = np.fft.fft(z, axis=1)
        if dealias == '3/2-rule':
            f1 = np.concatenate([f1[:, :N[1]//2], f1[:, M[1]-(N[1]-N[1]//2):]], axis=1)
        owner = L.forward_destinations()
        send = [torch.from_numpy(np.ascontiguousarray(f1[:, owner == d])) for d in range(world)]
        recv = [torch.zeros((L.M0l, L.N1l, L.Nh), dtype=torch.complex128) for _ in range(world)]
        _a2a(dist, recv, send, rank, world)
        W3 = np.concatenate([t.numpy() for t in recv], axis=0)         # (M0, N1l, Nh): x0 chunks in rank order
        f0 = np.fft.fft(W3, axis=0)
        if dealias == '3/2-rule':
            f0 = np.concatenate([f0[:N[0]//2], f0[M[0]-(N[0]-N[0]//2):]], axis=0)
        ref_f = o._fwd_p(v)[:, L.k1_slice]
        e_f = np.linalg.norm(f0-ref_f)/np.linalg.norm(ref_f)
        ret[rank] = (float(e_b), float(e_f), L.spectral_shape(), L.physical_shape())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('dealias', ['2/3-rule', '3/2-rule', 'None'])
def test_slab_pipeline_two_ranks_gloo(dealias):
    import torch.multiprocessing as mp
    N = (16, 32, 16)
    port = 29610 + ['2/3-rule', '3/2-rule', 'None'].index(dealias)
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dealias, N, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(2):
        e_b, e_f, ss, ps = ret[r]
        assert e_b < 1e-13 and e_f < 1e-13, (r, e_b, e_f)
        assert ss == (16, 16, 9)
        assert ps == ((12, 48, 24) if dealias == '3/2-rule' else (8, 32, 16))


def test_slab_layout_matches_c_plan_rules():
    from spectraldns_b200.slab import SlabLayout
    # 2/3-rule at N1=1024 on 8 ranks: middle ranks own no surviving mode (B0 has nothing to do there)
    K1l = [SlabLayout((1024, 1024, 1024), 8, r).K1l for r in range(8)]
    L0 = SlabLayout((1024, 1024, 1024), 8, 0)
    assert sum(K1l) == L0.K1n == 2*so_cut(1024)+1
    assert K1l[0] == 128 and K1l[3] == 0 and K1l[4] == 0 and K1l[7] == 128
    offs = [SlabLayout((1024, 1024, 1024), 8, r).c1off for r in range(8)]
    assert offs[0] == 0 and offs[1] == 128 and offs[7] == L0.K1n - 128
    with pytest.raises(ValueError):
        SlabLayout((32, 30, 32), 4, 0)


def so_cut(n):
    return int(np.ceil(2./3.*(n//2+1))) - 1


def _ce_worker(rank, world, port, dealias, N, nchunk, ret):
    """Copy-engine exchange replayed on CPU: send slots, per-chunk strided 2-D copies (spectraldns_b200.slab's
    mirror of csrc/sdns_api.cu) delivered as gloo messages, must land every element where the direct slab
    transpose puts it."""
    import torch
    import torch.distributed as dist
    from spectraldns_b200.slab import SlabLayout, ExchangePlan
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        Ls = [SlabLayout(N, world, r, dealias) for r in range(world)]
        Xs = [ExchangePlan(L, nchunk) for L in Ls]
        L, X = Ls[rank], Xs[rank]
        nf = 2

        def val_b0(f, x0, c1g, c2):      # unique tag of B0 output element (field, global x0, global compact k1, k2)
            return ((f*L.M[0] + x0)*L.K1n + c1g)*L.K2n + c2 + 1.0

        # ---- B0: own part straight into W0, peers' parts into the send slots
        W0 = np.zeros(nf*L.M0l*L.K1n*X.K2p)
        send = np.zeros(world*X.b0_slot(nf))
        f, x0, c1, c2 = np.meshgrid(np.arange(nf), np.arange(L.M[0]), np.arange(L.K1l), np.arange(L.K2n), indexing='ij')
        dest, x0l = x0//L.M0l, x0 % L.M0l
        v = val_b0(f, x0, c1 + L.c1off, c2)
        own = dest == rank
        W0[((f[own]*L.M0l + x0l[own])*L.K1n + L.c1off + c1[own])*X.K2p + c2[own]] = v[own]
        send[X.b0_send_index(nf, dest[~own], f[~own], x0l[~own], c1[~own], c2[~own])] = v[~own]
        reqs, inbox = [], []
        for c in range(nchunk):
            for k in range(1, world):
                r = (rank + k) % world
                cp = X.b0_copy(nf, r, c)
                if cp.width and cp.height:
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(cp.gather(send))), r, tag=c))
                s = (rank - k) % world                      # the copy rank s aims at this rank, same chunk
                cq = Xs[s].b0_copy(nf, rank, c)
                if cq.width and cq.height:
                    buf = torch.zeros((cq.height, cq.width), dtype=torch.float64)
                    reqs.append(dist.irecv(buf, s, tag=c))
                    inbox.append((cq, buf))
        for q in reqs:
            q.wait()
        for cq, buf in inbox:
            cq.scatter(W0, buf.numpy())
        W0 = W0.reshape(nf, L.M0l, L.K1n, X.K2p)[..., :L.K2n]
        f, x0l, c1g, c2 = np.meshgrid(np.arange(nf), np.arange(L.M0l), np.arange(L.K1n), np.arange(L.K2n), indexing='ij')
        e_b = float(np.abs(W0 - val_b0(f, x0l + rank*L.M0l, c1g, c2)).max())

        # ---- F1: element (field, global k1, global x0, k2) -> W3 (nf, N1l, M0, Nhp) on the rank owning k1
        def val_f1(f, k1, x0, c2):
            return ((f*L.N[1] + k1)*L.M[0] + x0)*L.Nh + c2 + 1.0
        W3 = np.zeros(nf*L.N1l*L.M[0]*X.Nhp)
        send = np.zeros(world*X.f1_slot(nf))
        f, k1, x0l, c2 = np.meshgrid(np.arange(nf), np.arange(L.N[1]), np.arange(L.M0l), np.arange(L.Nh), indexing='ij')
        dest, k1l = k1//L.N1l, k1 % L.N1l
        v = val_f1(f, k1, x0l + rank*L.M0l, c2)
        own = dest == rank
        W3[((f[own]*L.N1l + k1l[own])*L.M[0] + rank*L.M0l + x0l[own])*X.Nhp + c2[own]] = v[own]
        send[X.f1_send_index(nf, dest[~own], f[~own], k1l[~own], x0l[~own], c2[~own])] = v[~own]
        reqs, inbox = [], []
        for c in range(nchunk):
            for k in range(1, world):
                r = (rank + k) % world
                cp = X.f1_copy(nf, r, c)
                if cp.width and cp.height:
                    reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(cp.gather(send))), r, tag=100 + c))
                s = (rank - k) % world
                cq = Xs[s].f1_copy(nf, rank, c)
                if cq.width and cq.height:
                    buf = torch.zeros((cq.height, cq.width), dtype=torch.float64)
                    reqs.append(dist.irecv(buf, s, tag=100 + c))
                    inbox.append((cq, buf))
        for q in reqs:
            q.wait()
        for cq, buf in inbox:
            cq.scatter(W3, buf.numpy())
        W3 = W3.reshape(nf, L.N1l, L.M[0], X.Nhp)[..., :L.Nh]
        f, k1l, x0, c2 = np.meshgrid(np.arange(nf), np.arange(L.N1l), np.arange(L.M[0]), np.arange(L.Nh), indexing='ij')
        e_f = float(np.abs(W3 - val_f1(f, k1l + rank*L.N1l, x0, c2)).max())
        ret[rank] = (e_b, e_f, [Ls[r].K1l for r in range(world)])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('cfg', [(2, '2/3-rule', (16, 32, 16), 4), (2, '3/2-rule', (16, 32, 16), 3),
                                 (4, '2/3-rule', (16, 64, 24), 4), (4, 'None', (8, 16, 8), 2)])
def test_copy_engine_exchange_layout_gloo(cfg):
    import torch.multiprocessing as mp
    world, dealias, N, nchunk = cfg
    port = 29640 + [2, 4].index(world)*4 + ['2/3-rule', '3/2-rule', 'None'].index(dealias)
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_ce_worker, args=(r, world, port, dealias, N, nchunk, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    for r in range(world):
        e_b, e_f, K1l = ret[r]
        assert e_b == 0.0 and e_f == 0.0, (r, e_b, e_f, K1l)


def test_chunk_bounds_taper_and_cover():
    from spectraldns_b200.slab import chunk_bound
    for n in (0, 1, 5, 43, 64, 128, 341):
        for nc in (1, 2, 3, 4, 6, 16):
            b = [chunk_bound(n, nc, c) for c in range(nc+1)]
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
    b = [chunk_bound(90, 4, c) for c in range(5)]
    assert [y - x for x, y in zip(b, b[1:])] == [30, 30, 20, 10]


def test_cyclic_k1_ownership_host_side(monkeypatch):
    """Plan(k1_layout='cyclic') / SDNS_K1_LAYOUT=cyclic: rank r owns the axis-1 modes [r::P].  The host mirror reports it
    through T.local_slice(True), and the local wavenumbers / Nyquist mask follow from that slice.  Under the 2/3 rule the
    kept modes are then spread evenly, while the reference's contiguous blocks leave the end ranks 1.5x the average."""
    from spectraldns_b200 import spaces
    N, P = (16, 32, 8), 4
    bases = [spaces.FunctionSpace(n, 'F') for n in N]
    kept = lambda k: np.count_nonzero(np.abs(k) < 2./3*(N[1]//2 + 1))
    counts = {}
    for layout in ('blocks', 'cyclic'):
        monkeypatch.setenv('SDNS_K1_LAYOUT', layout)
        seen, counts[layout] = [], []
        for r in range(P):
            monkeypatch.setattr(spaces, 'world', lambda r=r: (r, P, 0))
            T = spaces.TensorProductSpace(None, bases)
            sl = T.local_slice(True)
            assert sl[1] == (slice(r, N[1], P) if layout == 'cyclic' else slice(r*8, r*8 + 8))
            assert T.shape(True) == (16, 8, 5)
            k = T.local_wavenumbers()[1].ravel()
            assert np.array_equal(k, np.fft.fftfreq(N[1], 1./N[1])[sl[1]])
            assert T.get_mask_nyquist().shape == T.shape(True)
            seen.extend(range(N[1])[sl[1]])
            counts[layout].append(kept(k))
        assert sorted(seen) == list(range(N[1]))          # a partition of the axis
    assert max(counts['cyclic']) - min(counts['cyclic']) <= 1
    assert max(counts['blocks']) == 8 and min(counts['blocks']) < 4
