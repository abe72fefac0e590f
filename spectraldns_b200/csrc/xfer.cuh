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
    unsigned int stage;                     // bytes per ring stage (>= every batch's piece, multiple of 128)
    int nrings;                             // warps of a transfer CTA that drive a ring of their own
    unsigned int inflight;                  // > 0: pick nctas so that the rings of the launch hold about this many bytes
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

// The role.  `cta` of `ncta` transfer CTAs; `smem` is the launch's dynamic shared memory (x.ring_bytes of it).
// Work unit = one row segment (row, seg) of a batch, sent to every destination in turn (the rotation starts at a
// different destination per unit, so that at any moment the senders spread their traffic over all receivers).  Up to
// SDNS_XW warps of the CTA drive independent rings (lane 0 of each): one thread's issue rate, not the ring, limits a
// single ring.  All index arithmetic is 32-bit and incremental -- a 64-bit division per piece costs more than the
// piece's copy.
#define SDNS_XW 4
#define SDNS_XHDR 512        // per ring: 16 mbarriers, 16 destination pointers, 16 byte counts
__device__ __forceinline__ void xfer_role(const XferArgs& x, unsigned char* smem, int cta, int ncta) {
    if (x.tma) {
        const int nwarp = (int)((blockDim.x + 31) >> 5);
        const int nw = nwarp < x.nrings ? nwarp : x.nrings;
        const int w = (int)(threadIdx.x >> 5);
        if ((threadIdx.x & 31) != 0 || w >= nw) return;
        const unsigned int sub = (x.ring_bytes / (unsigned int)nw) & ~127u;
        unsigned char* base = smem + (size_t)w * sub;
        unsigned long long* bars = reinterpret_cast<unsigned long long*>(base);
        char** mdst = reinterpret_cast<char**>(base + 128);
        unsigned int* mbytes = reinterpret_cast<unsigned int*>(base + 256);
        unsigned char* ring = base + SDNS_XHDR;
        const unsigned int stage = x.stage;
        int S = (int)((sub - SDNS_XHDR) / stage);
        if (S > 16) S = 16;
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_init_fence();
        const int L = S - 2;                                   // host guarantees S >= 3
        const unsigned int vc = (unsigned int)(cta * nw + w), nvc = (unsigned int)(ncta * nw);
        // producer cursor
        int bi = 0;
        unsigned int u = vc, d = 0, rot = 0, row = 0, off = 0, len = 0;
        bool have = false;
        auto seek = [&]() {                                    // position on unit u of batch bi (or run out of batches)
            while (bi < x.nbatch) {
                const XferBatch& b = x.b[bi];
                const unsigned int units = b.nrows * b.nseg;
                if (u < units) {
                    const unsigned int r = u / b.nseg, sg = u - r * b.nseg;
                    row = b.row0 + r; off = sg * b.piece;
                    len = b.width - off < b.piece ? b.width - off : b.piece;
                    rot = u % (unsigned int)b.ndest; d = 0; have = true;
                    return;
                }
                ++bi; u = vc;                                  // every batch is dealt out from its first unit
            }
            have = false;
        };
        seek();
        unsigned int k = 0, j = 0;                             // pieces produced / consumed
        while (have || j < k) {
            if (have) {
                const XferBatch& b = x.b[bi];
                const int s = (int)(k % (unsigned int)S);
                if (k >= (unsigned int)S) bulk_wait_read<1>();  // the store that last used this stage has read it
                unsigned int dd = d + rot; if (dd >= (unsigned int)b.ndest) dd -= b.ndest;
                const char* src = b.src[dd] + (unsigned long long)row * b.spitch + off;
                mdst[s] = b.dst[dd] + (unsigned long long)row * b.dpitch + off;
                mbytes[s] = len;
                mbar_expect_tx(&bars[s], len);
                bulk_g2s(ring + (size_t)s * stage, src, len, &bars[s]);
                ++k;
                if (++d == (unsigned int)b.ndest) { u += nvc; seek(); }
            }
            if (k - j > (unsigned int)L || (!have && j < k)) {
                const int s = (int)(j % (unsigned int)S);
                mbar_wait(&bars[s], (j / (unsigned int)S) & 1u);
                bulk_s2g(mdst[s], ring + (size_t)s * stage, mbytes[s]);
                bulk_commit();
                ++j;
            }
        }
        bulk_wait_all<0>();                                    // every store has been performed before the CTA retires
    } else {
        // no TMA (SDNS_EXCHANGE=ldst, or a ring too small): 16-byte loads and peer stores by all threads of the CTA
        for (int bi = 0; bi < x.nbatch; ++bi) {
            const XferBatch& b = x.b[bi];
            const unsigned int units = b.nrows * b.nseg;
            for (unsigned int u = (unsigned int)cta; u < units; u += (unsigned int)ncta) {
                const unsigned int r = u / b.nseg, sg = u - r * b.nseg;
                const unsigned int off = sg * b.piece;
                const unsigned int len = b.width - off < b.piece ? b.width - off : b.piece;
                const unsigned int rot = u % (unsigned int)b.ndest;
                for (int d = 0; d < b.ndest; ++d) {
                    int dd = d + (int)rot; if (dd >= b.ndest) dd -= b.ndest;
                    const uint4* sp = reinterpret_cast<const uint4*>(b.src[dd] + (unsigned long long)(b.row0 + r) * b.spitch + off);
                    uint4* dp = reinterpret_cast<uint4*>(b.dst[dd] + (unsigned long long)(b.row0 + r) * b.dpitch + off);
                    for (unsigned int i = threadIdx.x; i < len / 16; i += blockDim.x) dp[i] = sp[i];
                }
            }
        }
    }
}

// Host side, just before the launch: the ring is the launch's dynamic shared memory, split over up to SDNS_XW warps;
// about six stages per ring when it is large enough (pieces of at most 16 KB), at least three, else the role falls
// back to plain loads / stores.
inline void xfer_prepare(XferArgs& x, size_t smem, int threads) {
    if (x.nctas <= 0 || x.nbatch <= 0) { x.nctas = 0; x.nbatch = 0; return; }
    x.ring_bytes = (unsigned int)smem;
    // a ring is latency-bound (its stages in flight per round trip), so the transfer rate of a launch follows the
    // total ring capacity: small-CTA kernels get more transfer CTAs than the ones with 100 KB of shared memory
    if (x.inflight > 0 && smem > 0) {
        long long n = ((long long)x.inflight + (long long)smem - 1) / (long long)smem;
        x.nctas = (int)(n < 8 ? 8 : (n > 128 ? 128 : n));
    }
    int nw = threads / 32 < SDNS_XW ? threads / 32 : SDNS_XW;
    if (nw < 1) nw = 1;
    while (nw > 1 && ((smem / nw) & ~(size_t)127) < SDNS_XHDR + 6 * 1024) --nw;     // few large rings rather than many small ones
    x.nrings = nw;
    const size_t sub = (smem / nw) & ~(size_t)127;
    unsigned int piece = 0;
    if (sub >= SDNS_XHDR + 3 * 128) {
        piece = (unsigned int)((sub - SDNS_XHDR) / 6) & ~127u;
        if (piece < 128) piece = 128;
        if (piece > 16384) piece = 16384;
    }
    if (!x.tma || !piece) { x.tma = 0; piece = 16384; }
    unsigned int stage = 128;
    for (int i = 0; i < x.nbatch; ++i) {
        XferBatch& b = x.b[i];
        b.piece = b.width < piece ? b.width : piece;
        b.nseg = (b.width + b.piece - 1) / b.piece;
        stage = std::max(stage, (b.piece + 127u) & ~127u);
    }
    x.stage = stage;
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
