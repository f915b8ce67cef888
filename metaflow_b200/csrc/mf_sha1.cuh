// mf_sha1.cuh -- SHA-1 (FIPS 180-4) batch kernel for sm_100a.
//
// Replaces hashlib.sha1(blob).hexdigest() at metaflow/datastore/content_addressed_store.py:98.
// SHA-1 is a serial chain over the 64-byte blocks of ONE blob (Merkle-Damgard): the only
// parallelism is across blobs, and the speed of one stream is set by the dependent chain
//     a' = rol(a,5) + T,   T = f(b,c,d) + e + K + W          (one LEA.HI per round on the chain)
// so everything that is NOT on the chain is moved off the consumer's instruction stream:
//
//   unit = one consumer warp + two producer warps (even / odd blocks), 32 streams (one per lane);
//   CTA  = two units (256 threads; 216 KB of shared memory, so a CTA owns its SM: nothing else can sit on the
//          consumers' schedulers).  Warps 0,1 = consumers (sub-partitions 0,1, alone: warps 4,5 exit at once),
//          warps 2,6 / 3,7 = producers of unit 0 / 1 (sub-partitions 2 / 3)
//   producer: stages 512 bytes per stream with ONE TMA bulk copy per lane (cp.async.bulk + mbarrier
//             complete_tx; contiguous, one page per request -- a lane-per-blob 16-byte gather thrashes
//             the TLB and costs 25 instructions per 16 bytes), byte-swaps, expands
//             the message schedule W[0..79], adds K, and writes (K+W)[t][lane] to shared memory
//             (also builds the padded tail blocks: 0x80, zeros, 64-bit bit length);
//   consumer: per round only LDS(K+W) / LOP3(f) / IADD3(T) / LEA.HI(a') / SHF(rol30).
//   Producer b owns two (K+W) slots; hand-over through named barriers (bar.sync / bar.arrive, 64 threads).
// Blobs are sorted by length (host side) so the 32 streams of a CTA finish together.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

struct MfHashDesc {
    const uint8_t *src;   // device pointer, 16-byte aligned: the bytes to hash in THIS launch
    uint64_t len;         // their count; a multiple of 64 unless this is the last piece of the message
    uint32_t out_index;   // where the digest goes (last piece): digests + 20 * out_index
    uint32_t flags;       // MF_SHA_FIRST: start from the initial state; MF_SHA_LAST: pad and emit the digest
    uint64_t total_len;   // length of the whole message (for the padding of the last piece)
    uint32_t *state;      // 5 words carried between the pieces of one message (may be null when FIRST|LAST)
};
#define MF_SHA_FIRST 1u
#define MF_SHA_LAST 2u

__device__ __forceinline__ uint32_t mf_rol(uint32_t x, int s) { return __funnelshift_l(x, x, s); }
__device__ __forceinline__ uint32_t mf_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

#define MF_F_CH(b, c, d) ((d) ^ ((b) & ((c) ^ (d))))
#define MF_F_PAR(b, c, d) ((b) ^ (c) ^ (d))
#define MF_F_MAJ(b, c, d) (((b) & (c)) | ((d) & ((b) | (c))))

#define MF_SHA_GROUP 512u                 // bytes staged per stream per group (8 blocks)
#define MF_SHA_ROW (MF_SHA_GROUP + 16u)   // padded row: 16-byte reads of 8 consecutive lanes hit 8 distinct slots

#define MF_SHA_KWROW 84u                  // words per lane row of (K+W): 80 + 4 pad -> 16-byte accesses of 8 lanes are conflict-free

struct __align__(16) MfShaSmem {
    uint8_t raw[2][2][32][MF_SHA_ROW];    // [producer][group parity][stream][bytes]
    uint32_t kw[2][2][32][MF_SHA_KWROW];  // [producer][slot][stream][round]: two blocks of slack per producer
    unsigned long long mbar[2][2];        // [producer][group parity]: TMA bulk-copy completion barriers
    unsigned long long full[2][2];        // [producer][slot]: (K+W) slot filled (VARIANT 4: mbarrier instead of a named barrier)
};

// Named barriers: unit u uses ids 8u .. 8u+7 = FULL / EMPTY per (producer, slot).  A producer owns two (K+W)
// slots, so it may run two of its blocks ahead of the consumer.  The ids live in registers: the kernel uses all 16
// hardware barriers anyway, and one copy of the 80-round body stays resident in the instruction cache (unrolling
// the hand-over four times to get immediate ids made the consumer 1.5x slower: 25 KB of straight-line code).
__device__ __forceinline__ void mf_bar_sync_r(uint32_t id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void mf_bar_arrive_r(uint32_t id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
#define MF_BAR_FULL(unit, pid, slot) (8u * (unit) + 2u * (pid) + (slot))
#define MF_BAR_EMPTY(unit, pid, slot) (8u * (unit) + 4u + 2u * (pid) + (slot))

// ---- TMA (bulk async copy) + mbarrier helpers: one instruction per lane moves that lane's 512-byte group
__device__ __forceinline__ uint32_t mf_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mf_mbar_init(void *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mf_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mf_mbar_arrive_tx(void *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mf_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mf_mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mf_bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(mf_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(mf_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mf_mbar_wait(void *bar, uint32_t parity) {
    const uint32_t a = mf_smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > (1ll << 31)) __trap();       // ~1 s: never hang the GPU on a protocol bug
    }
}

// non-blocking test of a phase (the result can be looked at many instructions later), and the blocking wait
__device__ __forceinline__ uint32_t mf_mbar_test(void *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(mf_smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}

struct MfShaLane {               // what a lane knows about its stream in this launch
    const uint8_t *src; uint64_t len, nfull, nb, iters, bits; uint32_t r; bool valid, is_last, is_first;
};

// ------------------------------------------------------------------ producer PID of unit UNIT: blocks PID, PID+2, ...
template <bool MBAR>
__device__ __forceinline__ void mf_sha_produce(MfShaSmem &S, const MfShaLane &L, const uint32_t UNIT, const uint32_t PID, uint32_t lane, unsigned long long *dbg_out) {
    const uint8_t *src = L.src;
    const uint64_t len = L.len, iters = L.iters;
    const uint64_t ngroups = (iters + 7) >> 3;
    if (lane == 0) { mf_mbar_init(&S.mbar[PID][0], 32); mf_mbar_init(&S.mbar[PID][1], 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    // every lane moves ITS stream's next 512 bytes with one bulk copy (contiguous, one page per request)
    auto issue_group = [&](uint64_t g) {
        if (g >= ngroups) return;
        const uint64_t off = g * MF_SHA_GROUP;
        uint32_t nbytes = 0;
        if (src && off < len) { uint64_t rem = len - off; nbytes = rem >= MF_SHA_GROUP ? MF_SHA_GROUP : (uint32_t)((rem + 15) & ~15ull); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // earlier generic reads/writes of this row come first
        if (nbytes) { mf_mbar_arrive_tx(&S.mbar[PID][g & 1], nbytes); mf_bulk_g2s(&S.raw[PID][g & 1][lane][0], src + off, nbytes, &S.mbar[PID][g & 1]); }
        else mf_mbar_arrive(&S.mbar[PID][g & 1]);
    };
    auto wait_group = [&](uint64_t g) {
        mf_mbar_wait(&S.mbar[PID][g & 1], (uint32_t)((g >> 1) & 1));
        // zero what the copy did not bring: bytes past the end of the message (incl. the <= 15 over-read bytes)
        const uint64_t off = g * MF_SHA_GROUP;
        uint32_t have = 0;
        if (src && off < len) have = (len - off) >= MF_SHA_GROUP ? MF_SHA_GROUP : (uint32_t)(len - off);
        if (have < MF_SHA_GROUP) {
            uint8_t *row = &S.raw[PID][g & 1][lane][0];
            uint32_t k = have;
            for (; (k & 3u) && k < MF_SHA_GROUP; k++) row[k] = 0;
            for (; k < MF_SHA_GROUP; k += 4) *reinterpret_cast<uint32_t *>(row + k) = 0;
        }
    };
    issue_group(0);
    issue_group(1);
    long long pwait = 0, pstart = dbg_out ? clock64() : 0;
    uint64_t cur_group = ~0ull;
    for (uint64_t i = PID; i < iters; i += 2) {
        const uint32_t SLOT = (uint32_t)(i >> 1) & 1u;
        const uint64_t g = i >> 3;
        if (g != cur_group) { wait_group(g); cur_group = g; }
        uint32_t w[16];
        {
            const uint4 *row = reinterpret_cast<const uint4 *>(&S.raw[PID][g & 1][lane][(i & 7) * 64u]);
            uint4 v0 = row[0], v1 = row[1], v2 = row[2], v3 = row[3];
            w[0] = mf_bswap(v0.x); w[1] = mf_bswap(v0.y); w[2] = mf_bswap(v0.z); w[3] = mf_bswap(v0.w);
            w[4] = mf_bswap(v1.x); w[5] = mf_bswap(v1.y); w[6] = mf_bswap(v1.z); w[7] = mf_bswap(v1.w);
            w[8] = mf_bswap(v2.x); w[9] = mf_bswap(v2.y); w[10] = mf_bswap(v2.z); w[11] = mf_bswap(v2.w);
            w[12] = mf_bswap(v3.x); w[13] = mf_bswap(v3.y); w[14] = mf_bswap(v3.z); w[15] = mf_bswap(v3.w);
        }
        if (L.is_last && i >= L.nfull) {
            // padding: bytes past the message are zero in the staging row already
            if (i == L.nfull) {
                const uint32_t marker = 0x80u << (24 - 8 * (L.r & 3)), rk = L.r >> 2;
#pragma unroll
                for (int k = 0; k < 16; k++) w[k] |= marker & (0u - (uint32_t)(rk == (uint32_t)k));   // branch-free: keeps w[] in registers
            }
            if (i + 1 == L.nb) { w[14] = (uint32_t)(L.bits >> 32); w[15] = (uint32_t)L.bits; }
        }
        if (i >= 4) { long long t0_ = dbg_out ? clock64() : 0; mf_bar_sync_r(MF_BAR_EMPTY(UNIT, PID, SLOT)); if (dbg_out) pwait += clock64() - t0_; }
        uint4 *kw = reinterpret_cast<uint4 *>(&S.kw[PID][SLOT][lane][0]);
#define MF_W(t) (w[(t) & 15] = mf_rol(w[((t) + 13) & 15] ^ w[((t) + 8) & 15] ^ w[((t) + 2) & 15] ^ w[(t) & 15], 1))
#pragma unroll
        for (int t = 0; t < 16; t += 4) kw[t >> 2] = make_uint4(w[t] + 0x5A827999u, w[t + 1] + 0x5A827999u, w[t + 2] + 0x5A827999u, w[t + 3] + 0x5A827999u);
        { uint32_t a0 = MF_W(16), a1 = MF_W(17), a2 = MF_W(18), a3 = MF_W(19); kw[4] = make_uint4(a0 + 0x5A827999u, a1 + 0x5A827999u, a2 + 0x5A827999u, a3 + 0x5A827999u); }
#pragma unroll
        for (int t = 20; t < 40; t += 4) { uint32_t a0 = MF_W(t), a1 = MF_W(t + 1), a2 = MF_W(t + 2), a3 = MF_W(t + 3); kw[t >> 2] = make_uint4(a0 + 0x6ED9EBA1u, a1 + 0x6ED9EBA1u, a2 + 0x6ED9EBA1u, a3 + 0x6ED9EBA1u); }
#pragma unroll
        for (int t = 40; t < 60; t += 4) { uint32_t a0 = MF_W(t), a1 = MF_W(t + 1), a2 = MF_W(t + 2), a3 = MF_W(t + 3); kw[t >> 2] = make_uint4(a0 + 0x8F1BBCDCu, a1 + 0x8F1BBCDCu, a2 + 0x8F1BBCDCu, a3 + 0x8F1BBCDCu); }
#pragma unroll
        for (int t = 60; t < 80; t += 4) { uint32_t a0 = MF_W(t), a1 = MF_W(t + 1), a2 = MF_W(t + 2), a3 = MF_W(t + 3); kw[t >> 2] = make_uint4(a0 + 0xCA62C1D6u, a1 + 0xCA62C1D6u, a2 + 0xCA62C1D6u, a3 + 0xCA62C1D6u); }
#undef MF_W
        // no fence: bar.arrive orders this thread's prior shared-memory stores for the barrier's
        // participants (the PTX producer/consumer idiom); a MEMBAR here would drain every store
        if (MBAR) mf_mbar_arrive(&S.full[PID][SLOT]);      // release: this lane's stores above are visible to whoever sees the phase
        else mf_bar_arrive_r(MF_BAR_FULL(UNIT, PID, SLOT));
        // this producer is done with group g once it has taken its last block of it (blocks 6 or 7)
        if ((i & 7) >= 6) issue_group(g + 2);
    }
    if (dbg_out && lane == 0 && PID == 0) { dbg_out[0] = (unsigned long long)pwait; dbg_out[1] = (unsigned long long)(clock64() - pstart); }
}

// ------------------------------------------------------------------ consumer of unit UNIT
template <int VARIANT>
__device__ __forceinline__ void mf_sha_consume(MfShaSmem &S, const MfShaLane &L, const uint32_t UNIT, const MfHashDesc *desc, uint32_t lane,
                                               uint8_t *__restrict__ digests, unsigned long long *dbg_out, const uint32_t one, const uint32_t k30) {
    uint32_t h0 = 0x67452301u, h1 = 0xEFCDAB89u, h2 = 0x98BADCFEu, h3 = 0x10325476u, h4 = 0xC3D2E1F0u;
    if (L.valid && !L.is_first) {               // continue a message hashed piecewise (striped upload)
        const uint32_t *st = desc->state;
        h0 = st[0]; h1 = st[1]; h2 = st[2]; h3 = st[3]; h4 = st[4];
    }
    long long cwait = 0, cstart = dbg_out ? clock64() : 0;
    const uint64_t iters = L.iters, nb = L.nb;
    // VARIANT 4: the hand-over costs the lone consumer warp ~180 of its ~840 cycles per block (a blocking bar.sync, then the
    // first LDS of the new block before a round can start).  Here the FULL side is an mbarrier: the phase of block i+1 is
    // TESTED (non-blocking) 40 rounds before block i ends, looked at 20 rounds later, and the (K+W) words of the next
    // block's first 20 rounds are loaded under the last 20 rounds of this one.
    uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0, n4 = n0;
    auto full_wait = [&](uint64_t blk) {
        void *bar = &S.full[blk & 1u][(blk >> 1) & 1u];
        const uint32_t par = (uint32_t)(blk >> 2) & 1u;
        long long t0_ = dbg_out ? clock64() : 0;
        mf_mbar_wait(bar, par);
        if (dbg_out) cwait += clock64() - t0_;
    };
    if (VARIANT == 4 && iters) {
        full_wait(0);
        const uint4 *k = reinterpret_cast<const uint4 *>(&S.kw[0][0][lane][0]);
        n0 = k[0]; n1 = k[1]; n2 = k[2]; n3 = k[3]; n4 = k[4];
    }
    for (uint64_t i = 0; i < iters; i++) {
        const uint32_t PID = (uint32_t)i & 1u, SLOT = (uint32_t)(i >> 1) & 1u;
        if (VARIANT != 4) { long long t0_ = dbg_out ? clock64() : 0; mf_bar_sync_r(MF_BAR_FULL(UNIT, PID, SLOT)); if (dbg_out) cwait += clock64() - t0_; }
        const uint4 *kwv = reinterpret_cast<const uint4 *>(&S.kw[PID][SLOT][lane][0]);
        uint32_t a = h0, b = h1, c = h2, d = h3, e = h4;
        // (Measured: building T with two FMA-pipe IMADs instead of one ALU IADD3 is slower, 1040 vs 890
        // cycles per block; the consumer is bound by issued instructions, 4.25 per round this way.)
        // T = f + (e + (K+W)) is at least one round old; the chain is a' = rol(a,5) + T (LEA.HI).
        // The additions are pinned in PTX: left to the compiler, "+ f" is re-associated behind the
        // rotate-add and lands on the serial chain (LEA.HI -> IADD per round instead of LEA.HI alone).
// VARIANT 0: T = IADD3(e, K+W, f) on the ALU pipe (4 ALU-pipe instructions per round: LOP3, IADD3, LEA.HI, SHF).
// VARIANT 1: T by two IMADs with an opaque multiplier `one` (= 1, a kernel argument, so ptxas cannot fold them back
//            into an IADD3): 3 ALU-pipe + 2 FMA-pipe instructions per round; the two pipes issue on alternate cycles.
// VARIANT 3: rol(b,30) as IMAD.WIDE (b * 2^30: low word b<<30, high word b>>2) + IMAD: 3 ALU + 2 FMA, rotate off the ALU pipe.
#define MF_RND(F, a, b, c, d, e, kw_) { uint32_t f_ = F(b, c, d); \
        if (VARIANT == 1) { \
            asm("{\n\t.reg .u32 t_, r_;\n\tmad.lo.u32 t_, %1, %5, %2;\n\tmad.lo.u32 t_, %3, %5, t_;\n\tshf.l.wrap.b32 r_, %4, %4, 5;\n\tadd.u32 %0, r_, t_;\n\t}" \
                : "=r"(e) : "r"(e), "r"(kw_), "r"(f_), "r"(a), "r"(one)); \
        } else { \
            asm("{\n\t.reg .u32 t_, r_;\n\tadd.u32 t_, %1, %2;\n\tadd.u32 t_, t_, %3;\n\tshf.l.wrap.b32 r_, %4, %4, 5;\n\tadd.u32 %0, r_, t_;\n\t}" \
                : "=r"(e) : "r"(e), "r"(kw_), "r"(f_), "r"(a)); \
        } \
        if (VARIANT == 3) { \
            asm("{\n\t.reg .u64 p_;\n\t.reg .u32 lo_, hi_;\n\tmul.wide.u32 p_, %1, %2;\n\tmov.b64 {lo_, hi_}, p_;\n\tmad.lo.u32 %0, lo_, %3, hi_;\n\t}" \
                : "=r"(b) : "r"(b), "r"(k30), "r"(one)); \
        } else b = mf_rol(b, 30); }
#define MF_RND20(F, q) { \
        uint4 k0 = kwv[5 * (q)], k1 = kwv[5 * (q) + 1], k2 = kwv[5 * (q) + 2], k3 = kwv[5 * (q) + 3], k4 = kwv[5 * (q) + 4]; \
        MF_RND(F, a, b, c, d, e, k0.x) MF_RND(F, e, a, b, c, d, k0.y) MF_RND(F, d, e, a, b, c, k0.z) MF_RND(F, c, d, e, a, b, k0.w) MF_RND(F, b, c, d, e, a, k1.x) \
        MF_RND(F, a, b, c, d, e, k1.y) MF_RND(F, e, a, b, c, d, k1.z) MF_RND(F, d, e, a, b, c, k1.w) MF_RND(F, c, d, e, a, b, k2.x) MF_RND(F, b, c, d, e, a, k2.y) \
        MF_RND(F, a, b, c, d, e, k2.z) MF_RND(F, e, a, b, c, d, k2.w) MF_RND(F, d, e, a, b, c, k3.x) MF_RND(F, c, d, e, a, b, k3.y) MF_RND(F, b, c, d, e, a, k3.z) \
        MF_RND(F, a, b, c, d, e, k3.w) MF_RND(F, e, a, b, c, d, k4.x) MF_RND(F, d, e, a, b, c, k4.y) MF_RND(F, c, d, e, a, b, k4.z) MF_RND(F, b, c, d, e, a, k4.w) }
#define MF_RND20K(F, k0, k1, k2, k3, k4) { \
        MF_RND(F, a, b, c, d, e, k0.x) MF_RND(F, e, a, b, c, d, k0.y) MF_RND(F, d, e, a, b, c, k0.z) MF_RND(F, c, d, e, a, b, k0.w) MF_RND(F, b, c, d, e, a, k1.x) \
        MF_RND(F, a, b, c, d, e, k1.y) MF_RND(F, e, a, b, c, d, k1.z) MF_RND(F, d, e, a, b, c, k1.w) MF_RND(F, c, d, e, a, b, k2.x) MF_RND(F, b, c, d, e, a, k2.y) \
        MF_RND(F, a, b, c, d, e, k2.z) MF_RND(F, e, a, b, c, d, k2.w) MF_RND(F, d, e, a, b, c, k3.x) MF_RND(F, c, d, e, a, b, k3.y) MF_RND(F, b, c, d, e, a, k3.z) \
        MF_RND(F, a, b, c, d, e, k3.w) MF_RND(F, e, a, b, c, d, k4.x) MF_RND(F, d, e, a, b, c, k4.y) MF_RND(F, c, d, e, a, b, k4.z) MF_RND(F, b, c, d, e, a, k4.w) }
        if (VARIANT == 4) {
            MF_RND20K(MF_F_CH, n0, n1, n2, n3, n4)
            MF_RND20(MF_F_PAR, 1)
            const bool more = i + 1 < iters;
            void *nbar = &S.full[(i + 1) & 1u][((i + 1) >> 1) & 1u];
            const uint32_t npar = (uint32_t)((i + 1) >> 2) & 1u;
            const uint32_t ok = more ? mf_mbar_test(nbar, npar) : 1u;
            MF_RND20(MF_F_MAJ, 2)
            if (more) {
                if (!ok) full_wait(i + 1);                  // (a successful test was the acquire: the loads below come after it)
                const uint4 *k = reinterpret_cast<const uint4 *>(&S.kw[(i + 1) & 1u][((i + 1) >> 1) & 1u][lane][0]);
                n0 = k[0]; n1 = k[1]; n2 = k[2]; n3 = k[3]; n4 = k[4];
            }
            MF_RND20(MF_F_PAR, 3)
        } else {
        MF_RND20(MF_F_CH, 0)
        MF_RND20(MF_F_PAR, 1)
        MF_RND20(MF_F_MAJ, 2)
        MF_RND20(MF_F_PAR, 3)
        }
#undef MF_RND20K
#undef MF_RND20
#undef MF_RND
        if (i < nb) { h0 += a; h1 += b; h2 += c; h3 += d; h4 += e; }
        if (i + 4 < iters) mf_bar_arrive_r(MF_BAR_EMPTY(UNIT, PID, SLOT));
    }
    if (dbg_out && lane == 0) { dbg_out[2] = (unsigned long long)cwait; dbg_out[3] = (unsigned long long)(clock64() - cstart); }
    if (L.valid && !L.is_last) {
        uint32_t *st = desc->state;
        st[0] = h0; st[1] = h1; st[2] = h2; st[3] = h3; st[4] = h4;
    }
    if (L.valid && L.is_last) {
        uint8_t *out = digests + 20ull * desc->out_index;
        const uint32_t hh[5] = {h0, h1, h2, h3, h4};
#pragma unroll
        for (int k = 0; k < 5; k++) {
            out[4 * k] = (uint8_t)(hh[k] >> 24); out[4 * k + 1] = (uint8_t)(hh[k] >> 16);
            out[4 * k + 2] = (uint8_t)(hh[k] >> 8); out[4 * k + 3] = (uint8_t)hh[k];
        }
    }
}

#define MF_SHA_UNITS 2u
#define MF_SHA_STREAMS_PER_CTA (32u * MF_SHA_UNITS)
#define MF_SHA_THREADS 256u

// digests: 20 bytes per blob in hashlib.sha1().digest() byte order.  dbg (optional): 4 counters per unit.
// gate (optional): incremented once per CTA as soon as it is resident (see pack_submit).
template <int VARIANT>
__global__ void __launch_bounds__(MF_SHA_THREADS, 1) mf_sha1_kernel(const MfHashDesc *__restrict__ blobs, uint32_t n, uint8_t *__restrict__ digests,
                                                                  unsigned long long *__restrict__ dbg, uint32_t *__restrict__ gate,
                                                                  const uint32_t one, const uint32_t k30) {
    extern __shared__ __align__(16) uint8_t mf_sha_smem_raw[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (gate && threadIdx.x == 0) atomicAdd(gate, 1u);
    if (VARIANT == 4) {
        if (threadIdx.x < 8) mf_mbar_init(&reinterpret_cast<MfShaSmem *>(mf_sha_smem_raw)[threadIdx.x >> 2].full[(threadIdx.x >> 1) & 1u][threadIdx.x & 1u], 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    if (warp == 4 || warp == 5) return;                      // keep the consumers' sub-partitions to themselves
    const uint32_t unit = warp & 1u;
    MfShaSmem &S = reinterpret_cast<MfShaSmem *>(mf_sha_smem_raw)[unit];
    const uint32_t s = blockIdx.x * MF_SHA_STREAMS_PER_CTA + unit * 32u + lane;
    MfShaLane L;
    L.valid = s < n;
    const MfHashDesc *desc = blobs + (L.valid ? s : 0);
    L.src = L.valid ? desc->src : nullptr;
    L.len = L.valid ? desc->len : 0;
    L.nfull = L.len >> 6;
    L.r = (uint32_t)(L.len & 63u);
    const uint32_t sflags = L.valid ? desc->flags : (MF_SHA_FIRST | MF_SHA_LAST);
    L.is_last = (sflags & MF_SHA_LAST) != 0;
    L.is_first = (sflags & MF_SHA_FIRST) != 0;
    // blocks to run in this launch: the full blocks of the piece, plus the padding block(s) on the last piece
    L.nb = L.valid ? (L.is_last ? L.nfull + (L.r < 56 ? 1 : 2) : L.nfull) : 0;
    L.bits = (L.valid ? desc->total_len : 0) << 3;
    uint64_t iters = L.nb;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, iters, d); iters = o > iters ? o : iters; }
    L.iters = iters;
    unsigned long long *dbg_out = dbg ? dbg + ((size_t)blockIdx.x * MF_SHA_UNITS + unit) * 4 : nullptr;
    if (warp < 2) mf_sha_consume<VARIANT>(S, L, unit, desc, lane, digests, dbg_out, one, k30);
    else mf_sha_produce<VARIANT == 4>(S, L, unit, warp < 4 ? 0u : 1u, lane, dbg_out);
}
