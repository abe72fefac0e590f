// xfer.cuh -- the global transposes of the slab decomposition as a ROLE inside the pass kernels (sm_100a).
//
// mpi4py-fft moves the transposed data with one MPI_Alltoallw per transform (Transfer objects behind shenfun's
// TensorProductSpace; in-tree analogue spectralDNS3D_short.py:50-62).  Here the pass in front of a transpose (B0,
// F1) leaves the parts owned by other GPUs in per-destination send slots, and the NEXT pass kernels of the same
// stream carry the transfer: the first `nctas` CTAs of their grid do not transform anything -- one thread of each
// streams its share of the slots through a shared-memory ring with bulk-async copies
//     cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes   (local HBM -> ring, mbarrier-tracked)
//     cp.async.bulk.global.shared::cta.bulk_group                         (ring -> the peer GPU's array over NVLink)
// while every other CTA of the launch runs the FFT pass.  The TMA unit moves the bytes; the role costs one warp's
// issue slots and the ring.  No host-issued copy, no extra stream, no event: ordering is stream order plus the
// flag barrier that already separates the passes of different ranks.
#pragma once
#include "fft_core.cuh"

namespace sdns {

#define SDNS_XB_MAX 2        // batches one launch can carry
#define SDNS_XD_MAX 7        // destinations per batch (8 ranks)

// One 2-D copy shape (rows [row0, row0 + nrows) of `height`-row arrays) replicated over the destinations.
struct XferBatch {
    const char* src[SDNS_XD_MAX];
    char* dst[SDNS_XD_MAX];
    unsigned long long spitch, dpitch;      // bytes between rows
    unsigned int width;                     // bytes per row
    unsigned int row0, nrows;
    unsigned int nseg, piece;               // pieces per row, bytes per piece (the last one of a row may be shorter)
    int ndest;
};
struct XferArgs {
    int nctas;                              // CTAs of this launch that run the transfer role (0: none)
    int nbatch;
    int tma;                                // 1: bulk-async copies through the ring; 0: 16-byte loads / stores by the whole CTA
    unsigned int ring_bytes;                // dynamic shared memory available to the role
    XferBatch b[SDNS_XB_MAX];
};

#ifndef SDNS_HOST_SHIM
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, unsigned int bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (any global address this process has mapped: local HBM or a peer GPU's memory)
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, unsigned int bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#else
// emulation: a bulk copy is a memcpy that completes at once; the ring / stage bookkeeping runs unchanged
inline void mbar_init(void* bar, unsigned int) { *reinterpret_cast<unsigned long long*>(bar) = 0; }
inline void mbar_init_fence() {}
inline void mbar_expect_tx(void*, unsigned int) {}
inline void mbar_wait(void*, unsigned int) {}
inline void bulk_g2s(void* d, const void* s, unsigned int n, void*) { std::memcpy(d, s, n); }
inline void bulk_s2g(void* d, const void* s, unsigned int n) { std::memcpy(d, s, n); }
inline void bulk_commit() {}
template <int N> inline void bulk_wait_read() {}
template <int N> inline void bulk_wait_all() {}
#endif

// piece g (a global index over the batches of this launch) -> addresses.  Within a batch consecutive pieces go to
// consecutive destinations, starting with the one after this rank, so that at any moment every sender spreads its
// traffic over all receivers.
struct XferPiece { const char* src; char* dst; unsigned int bytes; };
__device__ __forceinline__ unsigned long long xfer_total(const XferArgs& x) {
    unsigned long long n = 0;
    for (int i = 0; i < x.nbatch; ++i) n += (unsigned long long)x.b[i].ndest * x.b[i].nrows * x.b[i].nseg;
    return n;
}
__device__ __forceinline__ XferPiece xfer_piece(const XferArgs& x, unsigned long long g) {
    int i = 0;
    for (; i < x.nbatch - 1; ++i) {
        const unsigned long long n = (unsigned long long)x.b[i].ndest * x.b[i].nrows * x.b[i].nseg;
        if (g < n) break;
        g -= n;
    }
    const XferBatch& b = x.b[i];
    const unsigned int d = (unsigned int)(g % (unsigned int)b.ndest);
    const unsigned long long r = g / (unsigned int)b.ndest;
    const unsigned int seg = (unsigned int)(r % b.nseg);
    const unsigned long long row = b.row0 + r / b.nseg;
    const unsigned int off = seg * b.piece;
    XferPiece p;
    p.src = b.src[d] + row * b.spitch + off;
    p.dst = b.dst[d] + row * b.dpitch + off;
    p.bytes = b.width - off < b.piece ? b.width - off : b.piece;
    return p;
}

// The role.  `cta` of `ncta` transfer CTAs; `smem` is the launch's dynamic shared memory (x.ring_bytes of it).
__device__ __forceinline__ void xfer_role(const XferArgs& x, unsigned char* smem, int cta, int ncta) {
    const unsigned long long total = xfer_total(x);
    if (x.tma) {
        if (threadIdx.x != 0) return;
        // ring: S stages of `stage` bytes after S mbarriers; lookahead S-2 so that the bulk store issued two pieces
        // ago may still be reading its stage when the next load is issued
        unsigned int stage = 0;
        for (int i = 0; i < x.nbatch; ++i) stage = x.b[i].piece > stage ? x.b[i].piece : stage;
        stage = (stage + 127u) & ~127u;
        int S = (int)((x.ring_bytes - 128u) / stage);
        if (S > 16) S = 16;
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem);
        unsigned char* ring = smem + 128;
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
        const int L = S - 2;                                   // host guarantees S >= 3
        unsigned long long n = 0;                              // my pieces: g = cta + k * ncta
        if (total > (unsigned long long)cta) n = (total - cta + ncta - 1) / ncta;
        for (unsigned long long k = 0; k < n + L; ++k) {
            if (k < n) {
                const int s = (int)(k % S);
                if (k >= (unsigned long long)S) bulk_wait_read<1>();     // the store that last used this stage has read it
                const XferPiece p = xfer_piece(x, cta + k * ncta);
                mbar_expect_tx(&bars[s], p.bytes);
                bulk_g2s(ring + (size_t)s * stage, p.src, p.bytes, &bars[s]);
            }
            if (k >= (unsigned long long)L) {
                const unsigned long long j = k - L;
                const int s = (int)(j % S);
                mbar_wait(&bars[s], (unsigned int)((j / S) & 1));
                const XferPiece p = xfer_piece(x, cta + j * ncta);
                bulk_s2g(p.dst, ring + (size_t)s * stage, p.bytes);
                bulk_commit();
            }
        }
        bulk_wait_all<0>();                                    // every store has been performed before the CTA retires
    } else {
        // no TMA (unaligned rows, or SDNS_XFER=ldst): 16-byte loads and peer stores by all threads of the CTA
        for (unsigned long long g = cta; g < total; g += ncta) {
            const XferPiece p = xfer_piece(x, g);
            const uint4* s = reinterpret_cast<const uint4*>(p.src);
            uint4* d = reinterpret_cast<uint4*>(p.dst);
            for (unsigned int i = threadIdx.x; i < p.bytes / 16; i += blockDim.x) d[i] = s[i];
        }
    }
}

// Host side, just before the launch: the ring is the launch's dynamic shared memory; about eight stages when it is
// large enough (pieces of at most 16 KB), at least three, else the role falls back to plain loads / stores.
inline void xfer_prepare(XferArgs& x, size_t smem) {
    if (x.nctas <= 0 || x.nbatch <= 0) { x.nctas = 0; x.nbatch = 0; return; }
    x.ring_bytes = (unsigned int)smem;
    unsigned int piece = 0;
    if (smem >= 128 + 3 * 128) {
        piece = (unsigned int)((smem - 128) / 8) & ~127u;
        if (piece < 128) piece = 128;
        if (piece > 16384) piece = 16384;
    }
    if (!x.tma || !piece) { x.tma = 0; piece = 16384; }
    for (int i = 0; i < x.nbatch; ++i) {
        XferBatch& b = x.b[i];
        b.piece = b.width < piece ? b.width : piece;
        b.nseg = (b.width + b.piece - 1) / b.piece;
    }
}

// A launch made only of the role: what is still pending when a pass needs the transposed data.
struct XferOnlyArgs { XferArgs x; };                    // the kernel lives in sdns_api.cu

// The first a.x.nctas CTAs (of row blockIdx.y == 0) of a launch run the transfer role; the others see the grid
// without them.  Declares bx / gx, the CTA index and grid size of the pass proper.
#define SDNS_XFER_ROLE(a, smraw)                                                              \
    int bx = (int)blockIdx.x, gx = (int)gridDim.x;                                            \
    if ((a).x.nctas > 0) {                                                                    \
        if (bx < (a).x.nctas) { if (blockIdx.y == 0) xfer_role((a).x, smraw, bx, (a).x.nctas); return; } \
        bx -= (a).x.nctas; gx -= (a).x.nctas;                                                 \
    }

}  // namespace sdns
