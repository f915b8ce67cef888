// mf_deflate2.cuh -- the member-owned DEFLATE encoder (round 2): one persistent CTA (32 warps) per SM takes whole
// gzip members (<= 64 chunks of 32 KiB = one 2 MiB stripe) by ticket and writes their FINAL bytes -- header, index,
// chunks back to back at a running offset, trailer -- so there is no per-chunk scratch slot and no gather pass.
//
// Replaces ContentAddressedStore._pack_v1 (metaflow/datastore/content_addressed_store.py:211-216: gzip.GzipFile(level 3)
// -> zlib deflate + zlib.crc32).  Same container format as mf_deflate.cuh (mf_core.h: MF v1), so the decoders are
// unchanged; what changed is how a chunk is produced (profiles/r01: 1.73 warp instructions per input byte, 55 % issue
// active, shared-memory pipe saturated by bank-conflicted table lookups and histogram atomics):
//
//   load      the next chunk arrives by ONE TMA bulk copy (cp.async.bulk + mbarrier) while the current one is encoded
//   LZ77      warp w owns 1 KiB segment w (as before: hash probe + in-window probes, ballot-ordered greedy parse)
//   sweep     ONE pass over the segment out of registers (lane = 32 bytes): CRC-32 with conflict-free 5-bit-sliced
//             tables (7 x 32-entry tables: every lookup of a warp hits 32 distinct banks or broadcasts), a 1-in-4 SAMPLED
//             literal histogram (4x fewer shared-memory atomics), and the Huffman codes looked up ONCE per byte in a
//             lane-group-striped table, packed into a per-lane staging row -- no separate bit-count sweep
//   codes     the chunk is packed speculatively with the codes of the previous chunk; the sampled histogram then prices
//             them against the chunk's own entropy bound (same 1/128 rule as before) and only a rejected chunk is
//             re-packed with fresh codes.  Every literal always has a code (floor count), so a kept code is always valid.
//   place     segment bit lengths -> one warp scan -> every lane shifts its staged bits to their final position in the
//             chunk image (funnel shifts; only a lane's first and last word are shared, OR-ed atomically)
//   store     the image leaves with coalesced 16-byte stores at the member's running offset; the partial last vector
//             is carried into the next chunk's image
//   CRC-32    per-thread Horner across the chunks of the member (multiply by x^(8*32768): 7 more conflict-free lookups),
//             one GF(2) multiply per thread per MEMBER instead of one per 64-byte piece per chunk
#pragma once
#include <cuda_runtime.h>
#include "mf_core.h"
#include "mf_sha1.cuh"      // mbarrier / bulk-copy helpers
#include "mf_deflate.cuh"   // code construction, header writer, warp helpers, CRC tables

struct MfEnc2Member {
    const uint8_t *src;     // raw bytes of the member (16-byte aligned, readable up to the next multiple of 16)
    uint8_t *dst;           // where the member's bytes go (16-byte aligned)
    uint64_t dst_cap;
    uint32_t raw_len;       // <= MF_MEMBER_RAW
    uint32_t n_chunks;
    uint32_t unit;          // index of out_len[] / status[]
    uint32_t first_chunk;   // global chunk index (phase profile only)
};

#define MF2_THREADS 1024
#define MF2_WARPS 32
#define MF2_ROW 17u                          // staging words per lane: 31 literals x 14 bits + one match token fit in 16
#define MF2_STAGE_WORDS (32u * MF2_ROW)      // 2176 bytes per warp: the LZ hash table (2 KiB) first, the staged bits afterwards
#define MF2_LL_MAXBITS 14
#define MF2_IMG_BYTES (MF_CHUNK + 64u)
#define MF2_BIAS16 (16.f * 4.f * 0.7213475f)    // 16 (units) * sample weight * 1 / (2 ln 2), per literal symbol in use
#define MF2_ACCEPT_TOL 128.f                  // codes are kept while they cost at most 1/128 more than fresh ones could
#define MF2_HIST_STEP 4                       // 8 of a lane's 32 byte positions are sampled, weight 4
// The sampled positions must not alias with the element size of the data (a stride of 4 over float32 data sees one
// byte lane of every element only -- the uniformly distributed low mantissa byte): both sets hold every residue mod 8
// exactly once (so every residue mod 4 twice), and segments alternate between them.
#define MF2_SAMPLE_A ((1u << 0) | (1u << 5) | (1u << 10) | (1u << 15) | (1u << 17) | (1u << 22) | (1u << 27) | (1u << 28))
#define MF2_SAMPLE_B ((1u << 2) | (1u << 7) | (1u << 8) | (1u << 13) | (1u << 19) | (1u << 20) | (1u << 25) | (1u << 30))

// 5-bit sliced CRC tables (conflict-free in shared memory): g_mf_crc5[j][v] = raw CRC state after the 32-bit word
// (v << 5j) followed by nothing, i.e. the slice-by-4 step applied to that word; g_mf_crcM[j][v] = (v << 5j) * x^(8*32768)
__device__ uint32_t g_mf_crc5[7][32];
__device__ uint32_t g_mf_crcM[7][32];
__device__ uint32_t g_mf_piece32_mul[1024];   // x^(8*32*j) mod P: moves a 32-byte piece's raw CRC j pieces left

__global__ void mf_init_tables2_kernel() {
    const int t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t < 7 * 32) {
        const int j = t >> 5, v = t & 31;
        const uint32_t x = j == 6 ? ((uint32_t)(v & 3) << 30) : ((uint32_t)v << (5 * j));
        g_mf_crc5[j][v] = g_mf_crc_tab[3][x & 0xFFu] ^ g_mf_crc_tab[2][(x >> 8) & 0xFFu] ^ g_mf_crc_tab[1][(x >> 16) & 0xFFu] ^ g_mf_crc_tab[0][x >> 24];
        g_mf_crcM[j][v] = mf_gf2_mul(x, mf_xpow8(MF_CHUNK));
    }
    if (t < 1024) g_mf_piece32_mul[t] = mf_xpow8(32ull * (uint64_t)t);
}

struct __align__(16) MfEnc2Smem {
    uint8_t data[2][MF_CHUNK + 64];                 // TMA double buffer
    uint32_t img[MF2_IMG_BYTES / 4];                // the chunk being assembled: carried bytes | header | tokens  (or stored)
    uint32_t stage[MF2_WARPS][MF2_STAGE_WORDS];
    uint32_t match[MF_SEGS_PER_CHUNK][MF_MAXM];
    uint32_t lit_codes[256 * 8];                    // literal codes, striped over 8 lane groups: [symbol][lane & 7]
    uint32_t crc5[7][32];
    uint32_t crcM[7][32];
    uint32_t crc8[4][256 * 4];                      // CRCV 1: slice-by-4 byte tables, striped over 4 lane groups: [k][byte][lane & 3]
    uint32_t ll_hist[288];
    uint32_t d_hist[32];
    uint32_t ll_codes[288];
    uint32_t d_codes[32];
    uint8_t ll_lens[288];
    uint8_t d_lens[32];
    MfWarpHdrScratch whdr;
    MfWarpCodeScratch wdist;
    uint32_t hdr_cache[MF_HDR_CACHE_WORDS];         // the block header of the codes in use, bit for bit from bit 0
    uint32_t cache_hdr_bits;                        // 0: no codes yet
    float cache_red16;
    uint32_t nmatch[MF_SEGS_PER_CHUNK];
    uint32_t seg_bits[MF_SEGS_PER_CHUNK];
    uint32_t seg_off[MF_SEGS_PER_CHUNK];
    uint32_t idx_chunk[MF_ENC_GROUP][2];            // member index being accumulated: {comp_bytes, hdr_bits}
    uint16_t idx_seg[MF_ENC_GROUP][MF_SEGS_PER_CHUNK];
    uint4 carry;                                    // bytes of the last, partial 16-byte vector of the previous chunk
    unsigned long long mbar[2];
    uint32_t crc_part[MF2_WARPS];
    uint32_t ticket, total_bits, crc_tail;
    uint32_t dec_cost, dec_miss, dec_used, dec_a_ll, dec_b_ll, dec_a_d, dec_b_d;     // this chunk against the codes in use
};

static_assert(sizeof(MfEnc2Smem) <= 227u * 1024u, "MfEnc2Smem must fit the 227 KB of shared memory a CTA can have");

__device__ __forceinline__ uint32_t mf2_fold7(const uint32_t (*T)[32], uint32_t x) {
    return T[0][x & 31u] ^ T[1][(x >> 5) & 31u] ^ T[2][(x >> 10) & 31u] ^ T[3][(x >> 15) & 31u] ^
           T[4][(x >> 20) & 31u] ^ T[5][(x >> 25) & 31u] ^ T[6][x >> 30];
}

// One slice-by-4 step, x = crc ^ word.  CRCV 0: seven 5-bit slices, conflict-free (27 instructions per word);
// CRCV 1: four byte slices out of 4-way lane-striped tables (15 instructions, ~2-way bank conflicts).
template <int CRCV>
__device__ __forceinline__ uint32_t mf2_crc_step(const MfEnc2Smem &S, uint32_t x, uint32_t lane) {
    if (CRCV == 0) return mf2_fold7(S.crc5, x);
    const uint32_t *T = &S.crc8[0][0] + (lane & 3u);
    return T[3072u + ((x & 0xFFu) << 2)] ^ T[2048u + ((x >> 6) & 0x3FCu)] ^ T[1024u + ((x >> 14) & 0x3FCu)] ^ T[(x >> 22) & 0x3FCu];
}

// Per-lane staging writer: LSB-first bits into the lane's row of shared memory, starting at bit 0.
struct Mf2Stage {
    uint32_t *row;
    uint32_t wi, acc, nacc;
    __device__ __forceinline__ void init(uint32_t *r) { row = r; wi = 0; acc = 0; nacc = 0; }
    __device__ __forceinline__ void put(uint32_t v, uint32_t nbits) {       // nbits <= 32, v < 2^nbits
        acc |= v << nacc;
        uint32_t n2 = nacc + nbits;
        if (n2 >= 32u) {
            row[wi++] = acc;
            acc = __funnelshift_l(v, 0u, nacc);           // v >> (32 - nacc); 0 when nacc == 0
            n2 -= 32u;
        }
        nacc = n2;
    }
    __device__ __forceinline__ uint32_t finish() { row[wi] = acc; row[wi + 1] = 0; return (wi << 5) + nacc; }
};

// One lane's 32-byte span of a segment: literal histogram sample, token packing into the staging row.
// Returns the number of bits staged (0 when !PACK).  wd[] = the span's bytes.
template <bool PACK, bool HIST, uint32_t SAMPLE, uint32_t WEIGHT = MF2_HIST_STEP>
__device__ __forceinline__ uint32_t mf2_span(MfEnc2Smem &S, const uint32_t (&wd)[8], uint32_t lo, uint32_t hi, const uint32_t *matches,
                                             uint32_t nmatch, uint32_t lane, uint32_t *row) {
    if (lo >= hi) { if (PACK) { row[0] = 0; row[1] = 0; } return 0; }
    uint32_t a = 0, b = nmatch;
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (mf_match_pos(matches[mid]) < lo) a = mid + 1; else b = mid; }
    uint32_t mi = a, p = lo;
    if (mi > 0) { const uint32_t pm = matches[mi - 1]; const uint32_t e = mf_match_pos(pm) + mf_match_len(pm); if (e > p) p = e; }
    uint32_t next_pos = mi < nmatch ? mf_match_pos(matches[mi]) : 0xFFFFFFFFu;
    const uint32_t *lit = S.lit_codes + (lane & 7u);
    Mf2Stage w;
    if (PACK) w.init(row);
    if (p == lo && next_pos >= hi && hi == lo + 32u) {
        // literal-only span: no token bookkeeping; two literals are joined (<= 28 bits) before they reach the writer
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
            const uint32_t b0 = __byte_perm(wd[k >> 2], 0u, 0x4440u + (k & 3)), b1 = __byte_perm(wd[k >> 2], 0u, 0x4440u + ((k + 1) & 3));   // PRMT
            if (HIST && ((SAMPLE >> k) & 1u)) atomicAdd(&S.ll_hist[b0], WEIGHT);
            if (HIST && ((SAMPLE >> (k + 1)) & 1u)) atomicAdd(&S.ll_hist[b1], WEIGHT);
            if (PACK) {
                const uint32_t e0 = lit[b0 * 8u], e1 = lit[b1 * 8u];
                const uint32_t l0 = e0 >> 16, l1 = e1 >> 16;
                w.put((e0 & 0xFFFFu) | ((e1 & 0xFFFFu) << l0), l0 + l1);
            }
        }
        return PACK ? w.finish() : 0u;
    }
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const uint32_t pos = lo + (uint32_t)k;
        if (pos < hi && pos == p) {
            if (pos == next_pos) {
                const uint32_t m = matches[mi];
                const uint32_t len = mf_match_len(m), dist = mf_match_dist(m);
                if (PACK) {
                    uint32_t ls, lb, lv, ds, db, dv;
                    mf_len_sym(len, ls, lb, lv); mf_dist_sym(dist, ds, db, dv);
                    const uint32_t le = S.ll_codes[ls], de = S.d_codes[ds];
                    w.put((le & 0xFFFFu) | (lv << (le >> 16)), (le >> 16) + lb);       // <= 14 + 5 bits
                    w.put((de & 0xFFFFu) | (dv << (de >> 16)), (de >> 16) + db);       // <= 15 + 13 bits
                }
                p += len; mi++;
                next_pos = mi < nmatch ? mf_match_pos(matches[mi]) : 0xFFFFFFFFu;
            } else {
                const uint32_t byte = __byte_perm(wd[k >> 2], 0u, 0x4440u + (k & 3));
                if (HIST && ((SAMPLE >> k) & 1u)) atomicAdd(&S.ll_hist[byte], WEIGHT);
                if (PACK) { const uint32_t e = lit[byte * 8u]; w.put(e & 0xFFFFu, e >> 16); }
                p++;
            }
        }
    }
    return PACK ? w.finish() : 0u;
}

// LZ77 over one segment by one warp (the parse of mf_deflate.cuh pass 2, literal histogram removed: literals are
// sampled in the sweep).  ht = 1024 x u16 hash table; returns the number of matches written to ml[].
// The table is cleared for a member's first chunk only: afterwards it holds whatever the previous chunk left there
// (its staged bits -- the region is shared), which is a function of the member's own bytes, so the output stays a
// function of the blob alone; a stale entry is just a candidate that fails the "precedes p and the bytes match" test.
__device__ __forceinline__ uint32_t mf2_lz_segment(MfEnc2Smem &S, const uint8_t *data, uint32_t base, uint32_t seglen, uint16_t *ht,
                                                   uint32_t *ml, uint32_t lane, bool lz, uint32_t &dry, bool clear) {
    if (!lz) return 0;
    if (clear) for (uint32_t i = lane; i < (1u << MF_HASH_BITS) / 2; i += 32) reinterpret_cast<uint32_t *>(ht)[i] = 0;
    __syncwarp();
    uint32_t nm = 0, pos = 0;
    while (pos < seglen) {
        const uint32_t p = pos + lane;
        const bool valid = p < seglen;
        // incompressible stretches: the longer nothing matched, the fewer windows are probed (every 4th after 3 dry
        // windows, every 8th after 6, every 16th after 12, one per segment after 48); `dry` is carried across the warp's
        // chunks of the member
        const uint32_t stride = dry >= 48u ? 1024u : dry >= 12u ? 512u : dry >= 6u ? 256u : 128u;
        if (dry >= 3u && (pos & (stride - 32u)) != 0) { pos = min(seglen, (pos + stride) & ~(stride - 1u)); continue; }
        const uint32_t nm_before = nm;
        const uint32_t w = valid ? mf_ld32u(data, base + p) : 0u;
        const bool can = (p + MF_MIN_MATCH <= seglen) && (nm < MF_MAXM);
        const uint32_t h = (w * 2654435761u) >> (32 - MF_HASH_BITS);
        int q = -1;
        if (can) { uint32_t cnd = ht[h]; if (cnd && cnd <= p) q = (int)cnd - 1; }
        __syncwarp();
        if (can) ht[h] = (uint16_t)(p + 1);      // lanes with equal hashes write the same slot: whichever lands is a valid candidate
        __syncwarp();
        {
            uint32_t w8 = __shfl_up_sync(0xFFFFFFFFu, w, 8);
            uint32_t w4 = __shfl_up_sync(0xFFFFFFFFu, w, 4);
            uint32_t w2 = __shfl_up_sync(0xFFFFFFFFu, w, 2);
            uint32_t w1 = __shfl_up_sync(0xFFFFFFFFu, w, 1);
            if (can) {
                if (lane >= 8 && w8 == w) q = (int)p - 8;
                if (lane >= 4 && w4 == w) q = (int)p - 4;
                if (lane >= 2 && w2 == w) q = (int)p - 2;
                if (lane >= 1 && w1 == w) q = (int)p - 1;
            }
        }
        bool has = false;
        if (can && q >= 0) has = (mf_ld32u(data, base + (uint32_t)q) == w);
        uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
        uint32_t cur = 0, adv = 32;
        while (true) {
            uint32_t m = cur < 32 ? (bal & (0xFFFFFFFFu << cur)) : 0u;
            if (!m) { adv = 32; break; }
            const uint32_t L = __ffs(m) - 1;
            const uint32_t pL = pos + L;
            const uint32_t qL = (uint32_t)__shfl_sync(0xFFFFFFFFu, q, L);
            const uint32_t maxlen = min((uint32_t)MF_MAX_MATCH, seglen - pL);
            uint32_t mlen = maxlen;
            for (uint32_t off = 4; off < maxlen; off += 32) {
                uint32_t o = off + lane;
                bool ne = (o >= maxlen) || (data[base + pL + o] != data[base + qL + o]);
                uint32_t bm = __ballot_sync(0xFFFFFFFFu, ne);
                if (bm) { mlen = off + (__ffs(bm) - 1); break; }
            }
            const uint32_t dist = pL - qL;
            const bool accept = (mlen >= 6) || (mlen >= 4 && dist <= 32) || (mlen == 5 && dist <= 256);
            if (accept) {
                if (lane == 0) {
                    ml[nm] = mf_match_pack(pL, mlen, dist);
                    uint32_t ls, lb, lv, ds, db, dv;
                    mf_len_sym(mlen, ls, lb, lv); mf_dist_sym(dist, ds, db, dv);
                    atomicAdd(&S.ll_hist[ls], 1u);
                    atomicAdd(&S.d_hist[ds], 1u);
                }
                nm++;
                if (nm >= MF_MAXM) bal = 0;
                const uint32_t e = L + mlen;
                if (e >= 32) { adv = e; break; }
                cur = e;
            } else {
                cur = L + 1;
            }
        }
        dry = (nm == nm_before) ? dry + 1 : 0;
        pos += adv;
    }
    return nm;
}

// flags bit 0: MFCAS_FLAG_NO_LZ
template <int CRCV>
__global__ void __launch_bounds__(MF2_THREADS, 1)
mf_deflate2_kernel(const MfEnc2Member *__restrict__ members, uint32_t n_members, uint64_t *__restrict__ out_len, int32_t *__restrict__ status,
                   uint32_t flags, unsigned long long *__restrict__ phase_clk, uint32_t *__restrict__ ticket) {
    extern __shared__ __align__(16) uint8_t mf_smem_raw[];
    MfEnc2Smem &S = *reinterpret_cast<MfEnc2Smem *>(mf_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool lz = (flags & 1u) == 0;
    for (uint32_t i = tid; i < 7 * 32; i += MF2_THREADS) { (&S.crc5[0][0])[i] = (&g_mf_crc5[0][0])[i]; (&S.crcM[0][0])[i] = (&g_mf_crcM[0][0])[i]; }
    if (CRCV == 1) for (uint32_t i = tid; i < 4096; i += MF2_THREADS) (&S.crc8[0][0])[i] = g_mf_crc_tab[i >> 10][(i >> 2) & 255u];
    if (tid == 0) { mf_mbar_init(&S.mbar[0], 1); mf_mbar_init(&S.mbar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t par0 = 0, par1 = 0;                   // phase parity of the two load barriers (uniform across the CTA)

    for (;;) {
        __syncthreads();                           // everything of the previous member is done with shared memory
        if (tid == 0) S.ticket = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t mi = S.ticket;
        if (mi >= n_members) break;
        const MfEnc2Member mem = members[mi];
        const uint32_t nc = mem.n_chunks;
        const uint32_t hdr_bytes = mf_member_header_bytes(nc);
        uint64_t woff = hdr_bytes;                 // running byte offset inside the member's destination
        const bool fits = (uint64_t)hdr_bytes + (uint64_t)mem.raw_len + 5ull * nc + MF_TRAILER <= mem.dst_cap;
        uint32_t crcB = 0;                         // this thread's Horner accumulator over the full chunks (see header)
        uint32_t dry = 0;
        if (tid == 0) { S.cache_hdr_bits = 0; S.carry = make_uint4(0, 0, 0, 0); S.crc_tail = 0; }
        for (uint32_t i = tid; i < 288; i += MF2_THREADS) S.ll_hist[i] = 0;
        if (tid < 32) S.d_hist[tid] = 0;
        if (tid == 32) { S.dec_cost = 0; S.dec_miss = 0; S.dec_used = 0; S.dec_a_ll = 0; S.dec_b_ll = 0; S.dec_a_d = 0; S.dec_b_d = 0; }
        { uint4 *z = reinterpret_cast<uint4 *>(S.img); for (uint32_t i = tid; i < MF2_IMG_BYTES / 16; i += MF2_THREADS) z[i] = make_uint4(0, 0, 0, 0); }
        __syncthreads();
        // first chunk's load
        auto issue_load = [&](uint32_t ci) {
            const uint32_t len = min((uint32_t)MF_CHUNK, mem.raw_len - ci * MF_CHUNK);
            const uint32_t nbytes = (len + 15u) & ~15u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mf_mbar_arrive_tx(&S.mbar[ci & 1u], nbytes);
            mf_bulk_g2s(&S.data[ci & 1u][0], mem.src + (size_t)ci * MF_CHUNK, nbytes, &S.mbar[ci & 1u]);
        };
        if (fits && nc && tid == 0) issue_load(0);

        for (uint32_t ci = 0; fits && ci < nc; ci++) {
#define MF2_PHASE(k) do { if (phase_clk && tid == 0) phase_clk[((size_t)mem.first_chunk + ci) * 16 + (k)] = clock64(); } while (0)
            MF2_PHASE(0);
            const uint32_t b = ci & 1u;
            const uint32_t len = min((uint32_t)MF_CHUNK, mem.raw_len - ci * MF_CHUNK);
            const bool last = (ci + 1 == nc);
            // a chunk of at most one segment (the ragged end of a blob, or a tiny blob) is stored: a dynamic block's header
            // alone is ~90 bytes, and pricing / building codes for it costs more than encoding a full chunk
            const bool tiny = len <= MF_SEG;
            if (!last && tid == 0) issue_load(ci + 1);           // buffer b^1 was last read two barriers ago
            // (histograms and pricing sums were cleared behind the previous chunk's barriers: no barrier here for a full chunk)
            if (b == 0) { mf_mbar_wait(&S.mbar[0], par0); par0 ^= 1u; } else { mf_mbar_wait(&S.mbar[1], par1); par1 ^= 1u; }
            uint8_t *data = S.data[b];
            if (len < MF_CHUNK) {                                 // bytes past the end: nothing downstream may depend on stale ones
                if (tid < 64 && len + tid < MF_CHUNK + 64u) data[len + tid] = 0;
                __syncthreads();
            }
            MF2_PHASE(1);

            // ---- LZ77 + sweep of this warp's segment
            const uint32_t seg = warp, base = seg * MF_SEG;
            const uint32_t seglen = base < len ? min((uint32_t)MF_SEG, len - base) : 0u;
            uint32_t nm = 0;
            if (seglen && !tiny) nm = mf2_lz_segment(S, data, base, seglen, reinterpret_cast<uint16_t *>(S.stage[warp]), S.match[seg], lane, lz, dry, ci == 0);
            __syncwarp();
            const bool have_codes = S.cache_hdr_bits != 0;
            if (have_codes && nm) {
                // a match whose length or distance symbol has no code in the codes in use is dropped (its bytes go out as
                // literals, which always have codes): a rare symbol never forces fresh codes by itself
                uint32_t kept = 0, dropped = 0;
                for (uint32_t i0 = 0; i0 < nm; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    uint32_t m = 0;
                    bool ok = false;
                    if (i < nm) {
                        m = S.match[seg][i];
                        uint32_t ls, lb, lv, ds, db, dv;
                        mf_len_sym(mf_match_len(m), ls, lb, lv); mf_dist_sym(mf_match_dist(m), ds, db, dv);
                        ok = S.ll_lens[ls] != 0 && S.d_lens[ds] != 0;
                    }
                    const uint32_t keep = __ballot_sync(0xFFFFFFFFu, ok);
                    __syncwarp();
                    if (ok) S.match[seg][kept + __popc(keep & ((1u << lane) - 1u))] = m;
                    kept += __popc(keep);
                    dropped += min(32u, nm - i0) - __popc(keep);
                    __syncwarp();
                }
                nm = kept;
                if (lane == 0 && dropped) atomicAdd(&S.dec_miss, dropped);
            }
            MF2_PHASE(2);
            const uint32_t lo = lane * 32u, hi = min(lo + 32u, seglen);
            uint32_t wd[8];
            {
                const uint4 v0 = *reinterpret_cast<const uint4 *>(data + base + lo), v1 = *reinterpret_cast<const uint4 *>(data + base + lo + 16);
                wd[0] = v0.x; wd[1] = v0.y; wd[2] = v0.z; wd[3] = v0.w; wd[4] = v1.x; wd[5] = v1.y; wd[6] = v1.z; wd[7] = v1.w;
            }
            // CRC-32
            if (!last) {
                uint32_t c = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) c = mf2_crc_step<CRCV>(S, c ^ wd[k], lane);
                crcB = mf2_fold7(S.crcM, crcB) ^ c;
            } else {
                // the member's last chunk (any length): full 32-byte pieces by table, positioned among the full pieces
                // (the shift past the ragged tail, x^(8 * (len & 31)), is applied once, after the reduction); the tail
                // piece -- at most one lane of the CTA has it -- bytewise
                uint32_t c = 0;
                const uint32_t g0 = base + lo;                   // byte offset of this lane's span in the chunk
                if (g0 + 32u <= len) {
#pragma unroll
                    for (int k = 0; k < 8; k++) c = mf2_crc_step<CRCV>(S, c ^ wd[k], lane);
                    c = mf_gf2_mul(c, g_mf_piece32_mul[(len >> 5) - 1u - (g0 >> 5)]);
                } else if (g0 < len) {
                    uint32_t t = 0;
                    for (uint32_t q = g0; q < len; q++) t = g_mf_crc_tab[0][(t ^ data[q]) & 0xFFu] ^ (t >> 8);
                    S.crc_tail = t;
                }
                c = mf_warp_xor(c);
                if (lane == 0) S.crc_part[warp] = c;
            }
            uint32_t bits = 0;
            if (seglen && !tiny) {
                uint32_t *row = S.stage[warp] + lane * MF2_ROW;
                if (seg & 1u) {
                    if (have_codes) bits = mf2_span<true, true, MF2_SAMPLE_B>(S, wd, lo, hi, S.match[seg], nm, lane, row);
                    else mf2_span<false, true, MF2_SAMPLE_B>(S, wd, lo, hi, S.match[seg], nm, lane, nullptr);
                } else {
                    if (have_codes) bits = mf2_span<true, true, MF2_SAMPLE_A>(S, wd, lo, hi, S.match[seg], nm, lane, row);
                    else mf2_span<false, true, MF2_SAMPLE_A>(S, wd, lo, hi, S.match[seg], nm, lane, nullptr);
                }
            }
            uint32_t incl = mf_warp_incl_scan(bits, lane);
            uint32_t lane_off = incl - bits;
            if (lane == 31) { S.seg_bits[seg] = incl; S.nmatch[seg] = nm; }
            if (tid == 0) S.ll_hist[256] = 1;                    // end-of-block (zeroed before the barrier above; nobody else adds to it)
            MF2_PHASE(3);
            __syncthreads();                                     // ---- barrier 1: histograms, staged bits, segment lengths complete

            // the image is all zero (every thread cleared the vectors it stored): the carried bytes go in front
            if (tid == 0) reinterpret_cast<uint4 *>(S.img)[0] = S.carry;
            // ---- price the codes in use against this chunk's (sampled) statistics: symbols 0..285 on warps 0..8, distance
            // symbols on warp 9 (integer sums: the verdict must not depend on the order of the atomics); warp 10 meanwhile
            // scans the segment lengths as if the codes are kept
            if (warp < 10) {
                uint32_t h = 0, l = 0;
                if (warp < 9) { if (tid < 286) { h = S.ll_hist[tid]; l = S.ll_lens[tid]; } }
                else if (lane < 30u) { h = S.d_hist[lane]; l = S.d_lens[lane]; }
                const uint32_t used = __popc(__ballot_sync(0xFFFFFFFFu, h != 0 && warp < 8));     // literals in use (sampled)
                const uint32_t cost = mf_warp_sum(h * l), a = mf_warp_sum(h);
                const uint32_t bb = mf_warp_sum(h > 1 ? (uint32_t)((float)h * __log2f((float)h) * 16.f + 0.5f) : 0u);
                if (lane == 0) {
                    atomicAdd(&S.dec_cost, cost);
                    if (used) atomicAdd(&S.dec_used, used);
                    if (warp < 9) { atomicAdd(&S.dec_a_ll, a); atomicAdd(&S.dec_b_ll, bb); }
                    else { atomicAdd(&S.dec_a_d, a); atomicAdd(&S.dec_b_d, bb); }
                }
            } else if (warp == 10) {
                const uint32_t sb = S.seg_bits[lane];
                const uint32_t sc = mf_warp_incl_scan(sb, lane);
                S.seg_off[lane] = S.cache_hdr_bits + sc - sb;
                if (lane == 31) S.total_bits = S.cache_hdr_bits + sc;
            }
            __syncthreads();                                     // ---- barrier 2: sums, offsets
            MF2_PHASE(4);
            bool accept;
            {
                float e16 = 0.f;
                { const uint32_t a = S.dec_a_ll, bsum = S.dec_b_ll; if (a > 1) e16 += (float)a * __log2f((float)a) * 16.f - (float)bsum; }
                { const uint32_t a = S.dec_a_d, bsum = S.dec_b_d; if (a > 1) e16 += (float)a * __log2f((float)a) * 16.f - (float)bsum; }
                // the plug-in entropy of a SAMPLED histogram is biased low by (K-1) / (2 ln 2) bits per ... sample set
                // (K symbols in use); with sample weight 4 that is MF2_BIAS16 units of bits*16 per symbol in use.  The same
                // term is added where the codes' own redundancy is measured (below), so the two sides stay comparable.
                e16 += MF2_BIAS16 * (float)S.dec_used;
                e16 = fmaxf(e16, 0.f);
                const float nsym = (float)(S.dec_a_ll + S.dec_a_d);
                // (dec_miss = matches dropped for want of a code: tolerated while they are few)
                accept = tiny || (have_codes && S.dec_miss * 16u <= S.dec_a_d + 128u &&
                                  (float)S.dec_cost * 16.f <= (e16 + S.cache_red16 * nsym) * (1.f + 1.f / MF2_ACCEPT_TOL) + 1024.f);
            }
            if (!accept) {
                // ---- fresh codes, then pack again.  They are built from EXACT literal counts: codes fitted to the 1-in-4
                // sample carry ~0.3 % of extra redundancy (the sample's own noise), which is most of the 1/128 budget the
                // following chunks are judged by -- every 4th chunk was rejected.  (Length / distance / end-of-block
                // counts are exact already.)  Every literal keeps a code: floor count 1.
                if (tid < 256) S.ll_hist[tid] = 0;
                if (tid < MF_HDR_CACHE_WORDS) S.hdr_cache[tid] = 0;
                __syncthreads();
                if (seglen) mf2_span<false, true, 0xFFFFFFFFu, 1u>(S, wd, lo, hi, S.match[seg], nm, lane, nullptr);
                __syncthreads();
                if (tid < 256 && S.ll_hist[tid] == 0) S.ll_hist[tid] = 1;
                __syncthreads();
                if (warp == 0) mf_build_code_warp<9>(S.ll_hist, 286, MF2_LL_MAXBITS, S.ll_lens, S.ll_codes, S.whdr.code, lane);
                else if (warp == 1) mf_build_code_warp<1>(S.d_hist, 30, 15, S.d_lens, S.d_codes, S.wdist, lane);
                __syncthreads();
                if (warp == 0) {
                    const uint32_t hb = mf_header_warp(S.ll_lens, S.d_lens, S.hdr_cache, S.whdr, lane, nullptr);
                    // redundancy of these codes on their own chunk, for the next chunks' verdicts
                    uint32_t own = 0, a = 0, bsum = 0;
                    for (uint32_t k = lane; k < 286u; k += 32) { const uint32_t h = S.ll_hist[k]; own += h * S.ll_lens[k]; a += h; bsum += h > 1 ? (uint32_t)((float)h * __log2f((float)h) * 16.f + 0.5f) : 0u; }
                    uint32_t ad = 0, bd = 0;
                    if (lane < 30) { const uint32_t h = S.d_hist[lane]; own += h * S.d_lens[lane]; ad = h; bd = h > 1 ? (uint32_t)((float)h * __log2f((float)h) * 16.f + 0.5f) : 0u; }
                    own = mf_warp_sum(own); a = mf_warp_sum(a); bsum = mf_warp_sum(bsum); ad = mf_warp_sum(ad); bd = mf_warp_sum(bd);
                    float e16 = 0.f;
                    if (a > 1) e16 += (float)a * __log2f((float)a) * 16.f - (float)bsum;
                    if (ad > 1) e16 += (float)ad * __log2f((float)ad) * 16.f - (float)bd;
                    e16 += 0.25f * MF2_BIAS16 * (float)S.dec_used;              // exact counts: weight 1 instead of 4
                    e16 = fmaxf(e16, 0.f);
                    const float nsym = (float)(a + ad);
                    if (lane == 0) { S.cache_hdr_bits = hb; S.cache_red16 = nsym > 0.f ? fmaxf((float)own * 16.f - e16, 0.f) / nsym : 0.f; }
                }
                for (uint32_t i = tid; i < 256u * 8u; i += MF2_THREADS) S.lit_codes[i] = S.ll_codes[i >> 3];
                __syncthreads();
                bits = 0;
                if (seglen) bits = mf2_span<true, false, 0u>(S, wd, lo, hi, S.match[seg], nm, lane, S.stage[warp] + lane * MF2_ROW);
                incl = mf_warp_incl_scan(bits, lane);
                lane_off = incl - bits;
                if (lane == 31) S.seg_bits[seg] = incl;
                __syncthreads();
                if (warp == 10) {
                    const uint32_t sb = S.seg_bits[lane];
                    const uint32_t sc = mf_warp_incl_scan(sb, lane);
                    S.seg_off[lane] = S.cache_hdr_bits + sc - sb;
                    if (lane == 31) S.total_bits = S.cache_hdr_bits + sc;
                }
                __syncthreads();
            }
            MF2_PHASE(5);
            // the histograms are not read again in this chunk: clear them for the next one (its LZ pass starts behind
            // barrier 4 below)
            for (uint32_t i = tid; i < 288; i += MF2_THREADS) S.ll_hist[i] = 0;
            if (tid < 32) S.d_hist[tid] = 0;
            // ---- chunk size, stored-or-not (every thread computes the same), index entry
            const uint32_t hdr_bits = S.cache_hdr_bits;
            bool stored;
            uint32_t comp_bytes;
            {
                const uint32_t eob = S.ll_codes[256];
                const uint32_t end_bits = S.total_bits + (eob >> 16) + 3u;          // + empty stored block header (000)
                const uint32_t dyn_bytes = ((end_bits + 7u) >> 3) + 4u;             // + LEN/NLEN of the sync marker
                stored = tiny || dyn_bytes >= len + 5u;
                comp_bytes = stored ? len + 5u : dyn_bytes;
            }
            if (warp == 11) {
                if (lane == 0) { S.idx_chunk[ci][0] = comp_bytes; S.idx_chunk[ci][1] = stored ? (uint32_t)MF_STORED_MARK : hdr_bits; }
                S.idx_seg[ci][lane] = stored ? (uint16_t)0 : (uint16_t)S.seg_bits[lane];
            }
            const uint32_t cbits = 8u * (uint32_t)(woff & 15u);   // the chunk starts behind the carried bytes
            if (!stored) {
                // every lane shifts its staged bits to their place: only its first and last word are shared
                if (bits) {
                    const uint32_t d0 = cbits + S.seg_off[seg] + lane_off;
                    const uint32_t sh = d0 & 31u, wb = d0 >> 5;
                    const uint32_t *row = S.stage[warp] + lane * MF2_ROW;
                    const uint32_t nout = (bits + sh + 31u) >> 5;                      // image words that get bits of this lane
                    uint32_t prev = 0;
                    for (uint32_t k = 0; k < nout; k++) {
                        const uint32_t cur = row[k];                                   // (one zero word follows the staged ones)
                        const uint32_t v = __funnelshift_l(prev, cur, sh);             // (cur << sh) | (prev >> (32 - sh))
                        if (k == 0 || k + 1u == nout) atomicOr(&S.img[wb + k], v);     // shared with the neighbouring lanes
                        else S.img[wb + k] = v;
                        prev = cur;
                    }
                }
                if (warp == 1) {                                  // the block header, bit for bit
                    const uint32_t sh = cbits & 31u, wb = cbits >> 5;
                    for (uint32_t k = lane; k < ((hdr_bits + 31u) >> 5); k += 32) {
                        const uint32_t v = S.hdr_cache[k];
                        atomicOr(&S.img[wb + k], v << sh);
                        if (sh) atomicOr(&S.img[wb + k + 1], v >> (32u - sh));
                    }
                }
                if (tid == 64) {                                  // end-of-block + sync marker (00 00 FF FF after bit padding)
                    const uint32_t eob = S.ll_codes[256];
                    const uint32_t at = cbits + S.total_bits;
                    const uint32_t v = eob & 0xFFFFu, sh = at & 31u;
                    atomicOr(&S.img[at >> 5], v << sh);
                    if (sh) atomicOr(&S.img[(at >> 5) + 1], v >> (32u - sh));
                    const uint32_t end_bits = S.total_bits + (eob >> 16) + 3u;
                    const uint32_t byte = (cbits >> 3) + ((end_bits + 7u) >> 3) + 2u;
                    atomicOr(&S.img[byte >> 2], 0xFFu << ((byte & 3u) * 8u));
                    atomicOr(&S.img[(byte + 1) >> 2], 0xFFu << (((byte + 1) & 3u) * 8u));
                }
            } else {
                // stored block: 00 | LEN | NLEN | raw bytes behind the carried bytes.  Every lane still holds its 32 raw
                // bytes in registers: they go to their (byte-shifted) place like staged bits do.
                const uint32_t g0 = base + lo;
                const uint32_t nby = g0 < len ? min(32u, len - g0) : 0u;
                if (nby) {
                    const uint32_t d0 = cbits + 40u + 8u * g0, nbits = 8u * nby;
                    const uint32_t sh = d0 & 31u, wb = d0 >> 5;
                    uint32_t prev = 0;
#pragma unroll
                    for (int k = 0; k <= 8; k++) {
                        uint32_t cur = 0;
                        if (k < 8) {
                            const uint32_t vb = nby > 4u * (uint32_t)k ? min(4u, nby - 4u * (uint32_t)k) : 0u;
                            cur = vb >= 4u ? wd[k < 8 ? k : 0] : (wd[k < 8 ? k : 0] & ((1u << (8u * vb)) - 1u));
                        }
                        if (32u * (uint32_t)k < nbits + sh) {
                            const uint32_t v = __funnelshift_l(prev, cur, sh);
                            const bool whole = k > 0 && 32u * ((uint32_t)k + 1u) <= nbits + sh;
                            if (whole) S.img[wb + k] = v; else atomicOr(&S.img[wb + k], v);
                        }
                        prev = cur;
                    }
                }
                if (tid == 64) {
                    uint8_t *ib = reinterpret_cast<uint8_t *>(S.img) + (cbits >> 3);    // nobody else touches these 5 bytes' words' other bytes concurrently except by atomicOr
                    const uint32_t hb[5] = {0u, len & 0xFFu, (len >> 8) & 0xFFu, (~len) & 0xFFu, ((~len) >> 8) & 0xFFu};
                    for (uint32_t j = 0; j < 5; j++) {
                        const uint32_t at = (cbits >> 3) + j;
                        atomicOr(&S.img[at >> 2], hb[j] << (8u * (at & 3u)));
                    }
                    (void)ib;
                }
            }
            MF2_PHASE(6);
            __syncthreads();                                     // ---- barrier 4: image complete
            {
                const uint32_t total = (uint32_t)(woff & 15u) + comp_bytes;
                const uint32_t nv = total >> 4;
                uint4 *g = reinterpret_cast<uint4 *>(mem.dst + (woff & ~15ull));
                const uint4 *s = reinterpret_cast<const uint4 *>(S.img);
                uint4 *sz = reinterpret_cast<uint4 *>(S.img);
                for (uint32_t i = tid; i < nv; i += MF2_THREADS) { g[i] = s[i]; sz[i] = make_uint4(0, 0, 0, 0); }   // stored, then cleared by its owner
                if (tid == (nv & (MF2_THREADS - 1u))) { S.carry = s[nv]; sz[nv] = make_uint4(0, 0, 0, 0); }       // the partial last vector goes with the next chunk
                if (tid == ((nv + 1u) & (MF2_THREADS - 1u))) sz[nv + 1u] = make_uint4(0, 0, 0, 0);                  // (a spill-over word of the last token)
            }
            woff += comp_bytes;
            if (tid == 32) { S.dec_cost = 0; S.dec_miss = 0; S.dec_used = 0; S.dec_a_ll = 0; S.dec_b_ll = 0; S.dec_a_d = 0; S.dec_b_d = 0; }   // read before barrier 4
            MF2_PHASE(7);
            // no barrier here: the next chunk touches the image only behind its own barrier 1, and everything else it
            // writes before that (hash tables / staging rows, match lists, histograms) was last read before barrier 4
#undef MF2_PHASE
        }

        // ---- member CRC: sum over threads of (Horner accumulator) * x^(8 * bytes behind the thread's piece), shifted past the
        // last chunk, plus the last chunk's own CRC (S.crc_part from its sweep)
        uint32_t crc_raw = 0;
        if (fits && nc) {
            uint32_t lastc = 0;
            if (warp == 0) lastc = mf_warp_xor(S.crc_part[lane]);          // the last chunk's full pieces
            __syncthreads();
            uint32_t c = 0;
            if (nc > 1) c = mf_gf2_mul(crcB, g_mf_piece32_mul[1023u - tid]);
            c = mf_warp_xor(c);
            if (lane == 0) S.crc_part[warp] = c;
            __syncthreads();
            if (warp == 0) {
                const uint32_t full = mf_warp_xor(S.crc_part[lane]);
                const uint32_t last_len = mem.raw_len - (nc - 1u) * MF_CHUNK;
                if (lane == 0) {
                    const uint32_t r = last_len & 31u;
                    const uint32_t last_raw = (r ? mf_gf2_mul(lastc, mf_xpow8(r)) : lastc) ^ S.crc_tail;
                    crc_raw = (nc > 1 ? mf_gf2_mul(full, mf_xpow8(last_len)) : 0u) ^ last_raw;
                }
            }
        }
        // ---- header + index (words), trailer (bytes), lengths
        if (fits) {
            uint32_t *gw = reinterpret_cast<uint32_t *>(mem.dst);
            const uint32_t nwords = hdr_bytes >> 2;
            const uint32_t *ic = &S.idx_chunk[0][0];
            const uint32_t *is = reinterpret_cast<const uint32_t *>(&S.idx_seg[0][0]);
            for (uint32_t j = tid; j < nwords; j += MF2_THREADS) {
                uint32_t v;
                if (j == 0) v = 0x04088b1fu;                                          // 1f 8b 08 | FLG.FEXTRA
                else if (j == 1) v = 0;                                                // MTIME 0: deterministic output
                else if (j == 2) v = (255u << 8) | ((hdr_bytes - 12u) << 16);          // XFL 0, OS unknown, XLEN
                else if (j == 3) v = (uint32_t)'M' | ((uint32_t)'F' << 8) | ((hdr_bytes - 16u) << 16);
                else if (j == 4) v = MF_FORMAT_VERSION | (MF_CHUNK_LOG2 << 8) | (MF_SEG_LOG2 << 16);
                else if (j == 5) v = mem.raw_len;
                else if (j == 6) v = nc;
                else if (j == 7) v = 0;
                else if (j < 8u + 2u * nc) {
                    const uint32_t e = j - 8u;                                         // {comp_bytes u32} {hdr_bits u16, rsv u16}
                    v = (e & 1u) ? (ic[e] & 0xFFFFu) : ic[e];
                } else v = is[j - 8u - 2u * nc];
                gw[j] = v;
            }
            if (tid == 0) {
                // trailer behind the carried bytes: final empty stored block, CRC-32, ISIZE
                uint8_t t[32];
                const uint32_t rem = (uint32_t)(woff & 15u);
                const uint4 cv = S.carry;
                const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w};
                for (uint32_t k = 0; k < rem; k++) t[k] = (uint8_t)(cw[k >> 2] >> (8u * (k & 3u)));
                uint8_t *q = t + rem;
                q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0xFF; q[4] = 0xFF;
                const uint32_t crc = mf_crc_finish(crc_raw, mem.raw_len);
                for (int k = 0; k < 4; k++) { q[5 + k] = (uint8_t)(crc >> (8 * k)); q[9 + k] = (uint8_t)(mem.raw_len >> (8 * k)); }
                uint8_t *g = mem.dst + (woff & ~15ull);
                for (uint32_t k = 0; k < rem + MF_TRAILER; k++) g[k] = t[k];
                out_len[mem.unit] = woff + MF_TRAILER;
            }
        } else if (tid == 0) {
            out_len[mem.unit] = (uint64_t)hdr_bytes + (uint64_t)mem.raw_len + 5ull * nc + MF_TRAILER;     // the need (upper bound)
            status[mem.unit] = 1;                                                      // MFCAS_E_DST_TOO_SMALL
        }
    }
}

// ------------------------------------------------------------------ member compaction (contiguous output API only)
// mfcas_pack_batch / mfcas_pack_device promise ONE contiguous RFC 1952 file per blob; the encoder writes members into
// fixed slots.  Layout: one CTA per blob scans its members' lengths; copy: one CTA per 64 KiB piece of a member.
struct MfCompactBlob { uint8_t *dst; uint64_t dst_cap; uint32_t first_member, n_members; };

__global__ void __launch_bounds__(256)
mf_compact_layout_kernel(const MfCompactBlob *__restrict__ blobs, const uint64_t *__restrict__ member_len, const int32_t *__restrict__ member_status,
                         uint64_t *__restrict__ member_dst_off, uint64_t *__restrict__ dst_len, int32_t *__restrict__ status) {
    __shared__ unsigned long long s_scan[256];
    __shared__ unsigned long long s_base;
    __shared__ int s_bad;
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    const MfCompactBlob blob = blobs[b];
    if (tid == 0) { s_base = 0; s_bad = 0; }
    __syncthreads();
    for (uint32_t m0 = 0; m0 < blob.n_members; m0 += 256) {
        const uint32_t m = m0 + tid;
        const unsigned long long mine = m < blob.n_members ? member_len[blob.first_member + m] : 0ull;
        if (m < blob.n_members && member_status[blob.first_member + m] != 0) s_bad = 1;
        s_scan[tid] = mine;
        __syncthreads();
        for (int d = 1; d < 256; d <<= 1) {
            const unsigned long long v = (int)tid >= d ? s_scan[tid - d] : 0ull;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        if (m < blob.n_members) member_dst_off[blob.first_member + m] = s_base + s_scan[tid] - mine;
        __syncthreads();
        if (tid == 0) s_base += s_scan[255];
        __syncthreads();
    }
    if (tid == 0) {
        const bool fit = s_base <= blob.dst_cap && !s_bad;
        dst_len[b] = s_base;
        status[b] = fit ? 0 : 1;
    }
}

#define MF_COMPACT_TILE 65536u
__global__ void __launch_bounds__(256)
mf_compact_copy_kernel(const MfCompactBlob *__restrict__ blobs, const uint32_t *__restrict__ member_blob, const uint8_t *__restrict__ slots, uint64_t slot_stride,
                       const uint64_t *__restrict__ member_len, const uint64_t *__restrict__ member_dst_off, const int32_t *__restrict__ status,
                       uint32_t tiles_per_member) {
    const uint32_t m = blockIdx.x / tiles_per_member, tile = blockIdx.x % tiles_per_member;
    const uint32_t b = member_blob[m];
    if (status[b] != 0) return;
    const uint64_t mlen = member_len[m];
    const uint64_t t0 = (uint64_t)tile * MF_COMPACT_TILE;
    if (t0 >= mlen) return;
    const uint32_t n = (uint32_t)min((uint64_t)MF_COMPACT_TILE, mlen - t0);
    uint8_t *dst = blobs[b].dst + member_dst_off[m] + t0;
    const uint8_t *src = slots + (size_t)m * slot_stride + t0;       // 16-byte aligned (slot stride and tile are multiples of 256)
    const uint32_t tid = threadIdx.x;
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    if (tid < head) dst[tid] = src[tid];
    const uint32_t body = (n - head) >> 4;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src);
    uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
    const uint32_t sh = (head & 3u) * 8u, w0 = head >> 2;
    for (uint32_t v = tid; v < body; v += blockDim.x) {
        const uint32_t i = w0 + 4u * v;
        const uint32_t a = sw[i], bb = sw[i + 1], cc = sw[i + 2], d = sw[i + 3], e = sh ? sw[i + 4] : 0u;
        uint4 o;
        o.x = __funnelshift_r(a, bb, sh); o.y = __funnelshift_r(bb, cc, sh);
        o.z = __funnelshift_r(cc, d, sh); o.w = __funnelshift_r(d, e, sh);
        dv[v] = o;
    }
    const uint32_t done = head + (body << 4);
    if (tid < n - done) dst[done + tid] = src[done + tid];
}

// In-place compaction of a blob whose members sit in fixed slots of its own destination (device-resident extents):
// member k moves from dst + k * slot_stride to the end of member k-1.  One CTA per blob, members in order, tiles in
// ascending order (the destination is always below the source, so a forward copy never overwrites unread bytes).
__global__ void __launch_bounds__(1024)
mf_compact_inplace_kernel(uint8_t *const *__restrict__ dst, const uint64_t *__restrict__ member_base, const uint64_t *__restrict__ member_len,
                          uint64_t slot_stride, uint64_t *__restrict__ dst_len) {
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    uint8_t *base = dst[b];
    const uint64_t m0 = member_base[b], m1 = member_base[b + 1];
    uint64_t out = 0;
    for (uint64_t m = m0; m < m1; m++) {
        const uint64_t len = member_len[m];
        const uint8_t *src = base + (m - m0) * slot_stride;
        uint8_t *d = base + out;
        if (d != src) {
            for (uint64_t t = 0; t < len; t += 16384) {               // 1024 threads x 16 bytes, loaded before anything is stored
                const uint64_t o = t + (uint64_t)tid * 16u;
                uint4 v = make_uint4(0, 0, 0, 0);
                uint32_t nb = 0;
                if (o < len) { nb = (uint32_t)min((uint64_t)16, len - o); v = *reinterpret_cast<const uint4 *>(src + o); }
                __syncthreads();
                if (nb == 16 && (((uintptr_t)(d + o)) & 15u) == 0) *reinterpret_cast<uint4 *>(d + o) = v;
                else { const uint32_t w[4] = {v.x, v.y, v.z, v.w}; for (uint32_t k = 0; k < nb; k++) d[o + k] = (uint8_t)(w[k >> 2] >> (8u * (k & 3u))); }
                __syncthreads();
            }
        }
        out += len;
    }
    if (tid == 0) dst_len[b] = out;
}
