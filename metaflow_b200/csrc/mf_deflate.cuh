// mf_deflate.cuh -- chunk-parallel DEFLATE encoder + CRC-32 for sm_100a, and the gather pass
// that assembles RFC 1952 members.
//
// Replaces ContentAddressedStore._pack_v1 (metaflow/datastore/content_addressed_store.py:211-216:
// gzip.GzipFile(compresslevel=3).write -> zlib deflate + zlib.crc32, /usr/lib/python3.12/gzip.py
// :259-285, :302-314, :350-360).  Output is a different (chunked, indexed) but equally valid
// RFC 1952 byte stream: parity is defined on the decoded bytes.
//
// Persistent CTAs (16 warps) take groups of MF_ENC_GROUP consecutive chunks of a member by ticket; per 32 KiB chunk,
// staged in shared memory:
//   pass 0  coalesced 16-byte loads HBM -> smem
//   pass 2  LZ77: each warp walks its 1 KiB segments 32 positions at a time (hash probe + in-window
//           probes via warp shuffles, cooperative match extension, ballot-ordered greedy parse); in stretches
//           without matches only every 4th / 8th / 16th window is probed
//   pass 3a price the chunk with the previous chunk's code lengths against its own entropy bound: within 1/128
//           the codes and the block header are kept (then pass 3 is only CRC-32 + a header copy)
//   pass 3  prefix codes by Kraft-slack refinement, warp-synchronous in registers (warp 0 literal/length,
//           warp 1 distance), then the dynamic-block header with position-parallel run-length tokens;
//           CRC-32 of the chunk meanwhile (64-byte piece per lane, slice-by-4 tables in smem, GF(2) combine)
//   pass 4  per-lane bit counting out of registers, warp-shuffle prefix scans -> every token's bit offset
//   pass 5  bit packing into smem, then coalesced 16-byte stores smem -> HBM slot
// HBM traffic per chunk: raw bytes read once, packed bytes written once (+ the gather copy).
#pragma once
#include <cuda_runtime.h>
#include "mf_core.h"

struct MfEncMember {
    const uint8_t *src;     // device pointer to the member's raw bytes (16-byte aligned)
    uint32_t raw_len;       // <= MF_MEMBER_RAW
    uint32_t n_chunks;
    uint32_t first_chunk;   // global chunk index of chunk 0
    uint32_t blob;          // owning blob
};

struct MfEncBlob {
    uint8_t *dst;           // device pointer where the packed blob goes
    uint64_t dst_cap;
    uint32_t first_member;
    uint32_t n_members;
};

__device__ __forceinline__ uint32_t mf_ld32u(const uint8_t *base, uint32_t off) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base);
    uint32_t i = off >> 2;
    return __funnelshift_r(w[i], w[i + 1], (off & 3u) * 8u);
}

__device__ __forceinline__ uint32_t mf_warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if ((int)lane >= d) v += u;
    }
    return v;
}
__device__ __forceinline__ uint32_t mf_warp_xor(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

// slice-by-4 tables for the reflected CRC-32: T[k][b] = raw crc of byte b followed by k zero bytes
__device__ uint32_t g_mf_crc_tab[4][256];
// x^(8*64*j) mod P for j = 0..511: multiplier that moves a 64-byte piece's raw CRC j pieces left
__device__ uint32_t g_mf_piece_mul[512];

__global__ void mf_init_tables_kernel() {
    int t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t < 256) {
        uint32_t c = (uint32_t)t;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? MF_CRC_POLY : 0u);
        g_mf_crc_tab[0][t] = c;
        uint32_t c1 = c;
        for (int s = 1; s < 4; s++) {
            // append one zero byte: crc' = T0[crc & 0xFF] ^ (crc >> 8); T0 entry recomputed bitwise
            uint32_t lo = c1 & 0xFFu, x = lo;
            for (int k = 0; k < 8; k++) x = (x >> 1) ^ ((x & 1u) ? MF_CRC_POLY : 0u);
            c1 = x ^ (c1 >> 8);
            g_mf_crc_tab[s][t] = c1;
        }
    }
    if (t < 512) g_mf_piece_mul[t] = mf_xpow8(64ull * (uint64_t)t);
}

// raw (zero-init, no final xor) CRC of `n` bytes (n multiple of 4) at smem offset `off` (multiple of 4)
__device__ __forceinline__ uint32_t mf_crc_words(const uint8_t *s, uint32_t off, uint32_t n, const uint32_t (*T)[256]) {
    uint32_t crc = 0;
#define MF_CRC_STEP(word) { const uint32_t x = crc ^ (word); crc = T[3][x & 0xFFu] ^ T[2][(x >> 8) & 0xFFu] ^ T[1][(x >> 16) & 0xFFu] ^ T[0][x >> 24]; }
    if (n == 64u && (off & 15u) == 0) {
        // a whole 64-byte piece: four 16-byte loads (word loads at the 64-byte lane stride are 16-way bank conflicts)
        const uint4 *v = reinterpret_cast<const uint4 *>(s + off);
#pragma unroll
        for (int q = 0; q < 4; q++) { const uint4 u = v[q]; MF_CRC_STEP(u.x) MF_CRC_STEP(u.y) MF_CRC_STEP(u.z) MF_CRC_STEP(u.w) }
        return crc;
    }
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s + off);
    for (uint32_t i = 0; i < (n >> 2); i++) MF_CRC_STEP(w[i])
#undef MF_CRC_STEP
    return crc;
}

// ------------------------------------------------------------------ warp-synchronous code construction
// One warp builds a complete prefix code for up to 32*K symbols entirely in registers + a few shared
// counters: symbol s = k*32 + lane.  Same Kraft-slack refinement as mf_huff_build (mf_core.h) but with the
// greedy order replaced by 256 benefit/cost buckets (one scan per round instead of a sort) and an exact
// completion step, so no loop is longer than K and nothing waits on a CTA barrier.
struct MfWarpCodeScratch {
    uint32_t bcost[32];      // code space wanted per benefit/cost bucket
    uint32_t lcnt[16];       // codes per length
    uint32_t nxt[16];        // next canonical code per length
};

__device__ __forceinline__ uint32_t mf_warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}
__device__ __forceinline__ uint32_t mf_warp_min(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}

template <int K>
__device__ __forceinline__ void mf_build_code_warp(const uint32_t *hist, int n, int maxbits, uint8_t *lens, uint32_t *codes,
                                                   MfWarpCodeScratch &W, uint32_t lane) {
    uint32_t f[K], l[K];
    uint32_t used = 0, T = 0;
#pragma unroll
    for (int k = 0; k < K; k++) { int s = k * 32 + (int)lane; f[k] = s < n ? hist[s] : 0u; used += f[k] != 0; T += f[k]; }
    used = mf_warp_sum(used);
    if (used < 2) {                       // zlib forces at least two codes; so do we (symbols 0, 1, ...)
        uint32_t need = 2 - used;
        uint32_t z = __ballot_sync(0xFFFFFFFFu, f[0] == 0);
        uint32_t rank = __popc(z & ((1u << lane) - 1u));
        if (f[0] == 0 && rank < need) { f[0] = 1; T += 1; }
    }
    T = mf_warp_sum(T);
    const uint32_t full = 1u << maxbits;
    uint32_t mass = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        uint32_t lk = 0;
        if (f[k]) {
            int a = 31 - __clz((int)T), b = 31 - __clz((int)f[k]);
            lk = (uint32_t)(a - b);
            if ((f[k] << lk) < T) lk++;               // smallest l with f * 2^l >= T  (Shannon length)
            lk = min(max(lk, 1u), (uint32_t)maxbits);
            mass += full >> lk;
        }
        l[k] = lk;
    }
    mass = mf_warp_sum(mass);
    while (mass > full) {                 // clamping to maxbits over-subscribed the code space: lengthen the rarest
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < K; k++) if (l[k] && l[k] < (uint32_t)maxbits) best = min(best, (f[k] << 9) | (uint32_t)(k * 32 + (int)lane));
        best = mf_warp_min(best);
        if (best == 0xFFFFFFFFu) break;
        const uint32_t bs = best & 511u;
        uint32_t newl = 0;
#pragma unroll
        for (int k = 0; k < K; k++) if ((uint32_t)(k * 32 + (int)lane) == bs) { l[k]++; newl = l[k]; }
        newl = __shfl_sync(0xFFFFFFFFu, newl, bs & 31u);
        mass -= full >> newl;
    }
    uint32_t slack = full - mass;
    const float inv_t = 32.0f / (float)T;
    for (int round = 0; round < 3 && slack; round++) {
        W.bcost[lane] = 0;
        __syncwarp();
        uint32_t bk[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            bk[k] = 0xFFFFFFFFu;
            if (l[k] > 1) {
                uint32_t key = f[k] << l[k];                     // in [T, 2T) unless the length was clamped
                uint32_t b = key > T ? min(31u, (uint32_t)((float)(key - T) * inv_t)) : 0u;
                bk[k] = b;
                atomicAdd(&W.bcost[b], full >> l[k]);
            }
        }
        __syncwarp();
        // lane j owns bucket 31-j (descending benefit/cost); accept the longest prefix of buckets that fits
        uint32_t cum = W.bcost[31 - lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, cum, d); if ((int)lane >= d) cum += u; }
        const uint32_t rej = __ballot_sync(0xFFFFFFFFu, cum > slack);
        uint32_t thr = 0, consumed;
        if (rej) {
            const int L = __ffs(rej) - 1;                         // first bucket (in descending order) that does not fit
            thr = 32u - (uint32_t)L;
            consumed = L ? __shfl_sync(0xFFFFFFFFu, cum, L - 1) : 0u;
        } else {
            consumed = __shfl_sync(0xFFFFFFFFu, cum, 31);
        }
#pragma unroll
        for (int k = 0; k < K; k++) if (bk[k] != 0xFFFFFFFFu && bk[k] >= thr) l[k]--;
        if (consumed == 0) break;
        slack -= consumed;
        __syncwarp();
    }
    // exact completion: DEFLATE decoders reject an under-full code, the Kraft sum must end at exactly 1
    if (lane < 16) W.lcnt[lane] = 0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; k++) if (l[k]) atomicAdd(&W.lcnt[l[k]], 1u);
    __syncwarp();
    for (int guard = 0; slack && guard < 64; guard++) {
        const uint32_t myc = (lane >= 2 && (int)lane <= maxbits) ? W.lcnt[lane] : 0u;
        const uint32_t ok = __ballot_sync(0xFFFFFFFFu, myc > 0 && (full >> lane) <= slack);
        if (!ok) break;
        const uint32_t L = (uint32_t)__ffs(ok) - 1u;      // shortest such length = largest piece of code space
        const uint32_t cst = full >> L;
        const uint32_t cntL = __shfl_sync(0xFFFFFFFFu, myc, (int)L);
        const uint32_t take = min(cntL, slack / cst);
        uint32_t done = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const bool cand = l[k] == L;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, cand);
            const uint32_t rank = __popc(m & ((1u << lane) - 1u));
            if (cand && done + rank < take) l[k]--;
            done += __popc(m);
        }
        __syncwarp();
        if (lane == 0) { W.lcnt[L] -= take; W.lcnt[L - 1] += take; }
        slack -= take * cst;
        __syncwarp();
    }
    // canonical codes (RFC 1951 3.2.2): first code per length, then rank among equal lengths in symbol order
    if (lane < 16) {                      // first code of length `lane`: sum_{j<lane} cnt[j] << (lane - j)
        uint32_t code = 0;
        for (uint32_t j = 1; j < lane; j++) code += W.lcnt[j] << (lane - j);
        W.nxt[lane] = code;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t L = l[k];
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, L);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        const uint32_t code = L ? W.nxt[L] + rank : 0u;
        __syncwarp();
        if (L && rank == 0) W.nxt[L] += __popc(peers);
        __syncwarp();
        const int s = k * 32 + (int)lane;
        if (s < n) { lens[s] = (uint8_t)L; codes[s] = L ? ((L << 16) | (__brev(code) >> (32 - L))) : 0u; }
    }
    __syncwarp();
}

// Warp-synchronous dynamic-block header (RFC 1951 3.2.7): every position of the code-length sequence decides
// locally which run-length token (if any) starts there, the code-length code is built from the token
// histogram, and the tokens are placed with a prefix scan of their bit lengths.  out = zeroed words.
struct MfWarpHdrScratch {
    MfWarpCodeScratch code;
    uint32_t clhist[32];
    uint32_t clcodes[32];
    uint8_t cllens[32];
    uint8_t seq[320];
};

__device__ __forceinline__ uint32_t mf_header_warp(const uint8_t *ll_lens, const uint8_t *d_lens, uint32_t *out, MfWarpHdrScratch &H, uint32_t lane,
                                                   unsigned long long *ph) {
    // HLIT / HDIST: trim trailing zero lengths
    int hlit = 257, hdist = 1;
    for (int k = 8; k >= 0; k--) {
        int s = k * 32 + (int)lane;
        uint32_t m = __ballot_sync(0xFFFFFFFFu, s < 286 && ll_lens[s] != 0);
        if (m) { hlit = max(257, k * 32 + (32 - __clz((int)m))); break; }
    }
    {
        uint32_t m = __ballot_sync(0xFFFFFFFFu, lane < 30 && d_lens[lane] != 0);
        if (m) hdist = max(1, 32 - __clz((int)m));
    }
    const int n = hlit + hdist;
    for (int p = (int)lane; p < n; p += 32) H.seq[p] = p < hlit ? ll_lens[p] : d_lens[p - hlit];
    if (lane < 32) H.clhist[lane] = 0;
    __syncwarp();
    if (ph && lane == 0) ph[4] = clock64();
    // run start a(p) and run end e(p) from ballots of "a run starts here" flags (rows k = p / 32): no scan,
    // no carried dependency between rows except two uniform register chains
    uint32_t v[10], m[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        const int p = k * 32 + (int)lane;
        const uint32_t cur = p < n ? H.seq[p] : 0x100u;
        const uint32_t prev = (p > 0 && p <= n) ? H.seq[p - 1] : 0x200u;
        v[k] = cur;
        m[k] = __ballot_sync(0xFFFFFFFFu, p >= n || p == 0 || cur != prev);
    }
    if (ph && lane == 0) ph[5] = clock64();
    uint32_t a[10], e[10];
    {
        uint32_t last = 0;                                    // last start before row k
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const uint32_t below = m[k] & (0xFFFFFFFFu >> (31u - lane));          // starts at lanes <= mine
            a[k] = below ? (uint32_t)(k * 32 + 31 - __clz((int)below)) : last;
            if (m[k]) last = (uint32_t)(k * 32 + 31 - __clz((int)m[k]));
        }
        uint32_t nxt = 320u;                                  // first start after row k
#pragma unroll
        for (int k = 9; k >= 0; k--) {
            const uint32_t above = lane == 31 ? 0u : (m[k] & (0xFFFFFFFFu << (lane + 1u)));   // starts at lanes > mine
            e[k] = above ? (uint32_t)(k * 32 + __ffs((int)above) - 1) : nxt;
            if (m[k]) nxt = (uint32_t)(k * 32 + __ffs((int)m[k]) - 1);
        }
    }
    if (ph && lane == 0) ph[0] = clock64();
    // token decision per position
    uint32_t tsym[10], text[10];          // tsym 0xFF = no token here
#pragma unroll
    for (int k = 0; k < 10; k++) {
        const int p = k * 32 + (int)lane;
        uint32_t sym = 0xFFu, ext = 0;
        if (p < n) {
            const uint32_t R = e[k] - a[k], off = (uint32_t)p - a[k], val = v[k];
            if (val == 0) {
                const uint32_t fullr = R / 138u, rem = R - fullr * 138u;
                if (off < fullr * 138u) { if (off % 138u == 0) { sym = 18; ext = 127; } }
                else {
                    const uint32_t ro = off - fullr * 138u;
                    if (rem >= 11) { if (ro == 0) { sym = 18; ext = rem - 11; } }
                    else if (rem >= 3) { if (ro == 0) { sym = 17; ext = rem - 3; } }
                    else sym = 0;
                }
            } else if (off == 0) sym = val;
            else {
                const uint32_t rest = R - 1, o = off - 1, fullg = rest / 6u, rem = rest - fullg * 6u;
                if (o < fullg * 6u) { if (o % 6u == 0) { sym = 16; ext = 3; } }
                else {
                    const uint32_t ro = o - fullg * 6u;
                    if (rem >= 3) { if (ro == 0) { sym = 16; ext = rem - 3; } }
                    else sym = val;
                }
            }
            if (sym != 0xFFu) atomicAdd(&H.clhist[sym], 1u);
        }
        tsym[k] = sym; text[k] = ext;
    }
    __syncwarp();
    if (ph && lane == 0) ph[1] = clock64();
    mf_build_code_warp<1>(H.clhist, 19, 7, H.cllens, H.clcodes, H.code, lane);
    if (ph && lane == 0) ph[2] = clock64();
    // HCLEN
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t ord = 0;
#pragma unroll
    for (int i = 0; i < 19; i++) if ((int)lane == i) ord = order[i];
    const uint32_t nz = __ballot_sync(0xFFFFFFFFu, lane < 19 && H.cllens[ord] != 0);
    const int hclen = max(4, 32 - __clz((int)nz));
    // fixed part: BFINAL=0, BTYPE=10, HLIT, HDIST, HCLEN, then HCLEN x 3 bits
    if (lane == 0) {
        uint32_t w0 = 0u | (2u << 1) | ((uint32_t)(hlit - 257) << 3) | ((uint32_t)(hdist - 1) << 8) | ((uint32_t)(hclen - 4) << 13);
        atomicOr(&out[0], w0);
    }
    if ((int)lane < hclen) {
        const uint32_t bit = 17u + 3u * lane, val = H.cllens[ord];
        atomicOr(&out[bit >> 5], val << (bit & 31u));
        if ((bit & 31u) > 29u) atomicOr(&out[(bit >> 5) + 1], val >> (32u - (bit & 31u)));
    }
    // tokens: prefix scan of bit lengths in position order
    uint32_t base = 17u + 3u * (uint32_t)hclen;
#pragma unroll
    for (int k = 0; k < 10; k++) {
        uint32_t bits = 0, word = 0;
        if (tsym[k] != 0xFFu) {
            const uint32_t ce = H.clcodes[tsym[k]];
            const uint32_t cl = ce >> 16, eb = tsym[k] == 16 ? 2u : (tsym[k] == 17 ? 3u : (tsym[k] == 18 ? 7u : 0u));
            word = (ce & 0xFFFFu) | (text[k] << cl);
            bits = cl + eb;
        }
        uint32_t incl = bits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += u; }
        if (bits) {
            const uint32_t bit = base + incl - bits;
            atomicOr(&out[bit >> 5], word << (bit & 31u));
            if ((bit & 31u) + bits > 32u) atomicOr(&out[(bit >> 5) + 1], word >> (32u - (bit & 31u)));
        }
        base += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    return base;      // header length in bits
}

// 32-bit form of MfBitWriter for the device sweeps (same bits: LSB first, first word OR-ed, rest stored).
struct MfBitWriterDev {
    uint32_t *out;
    uint32_t wi, acc, nacc;
    bool first;
    __device__ __forceinline__ void init(uint32_t *o, uint32_t bitpos) { out = o; wi = bitpos >> 5; acc = 0; nacc = bitpos & 31u; first = true; }
    __device__ __forceinline__ void put(uint32_t v, uint32_t nbits) {      // nbits <= 32, v < 2^nbits
        acc |= v << nacc;
        uint32_t n2 = nacc + nbits;
        if (n2 >= 32u) {
            if (first) { atomicOr(&out[wi], acc); first = false; } else out[wi] = acc;
            wi++;
            acc = __funnelshift_l(v, 0u, nacc);          // the bits of v that did not fit: v >> (32 - nacc), 0 when nacc == 0
            n2 -= 32u;
        }
        nacc = n2;
    }
    __device__ __forceinline__ void flush() { if (nacc) atomicOr(&out[wi], acc); }
};

// Device form of mf_lane_emit (mf_core.h keeps the portable one for the CPU model): same tokens, same bits.
// The portable loop reads the segment a byte at a time at a 32-byte lane stride -- an 8-way bank conflict per
// read, twice per chunk, which made the two sweeps the shared-memory bottleneck of the encoder.  Here a lane
// takes its 32-byte span with two 16-byte loads and walks it out of registers (fully unrolled, static byte
// extraction); the counting sweep looks lengths up in the byte table (72 words: mostly broadcast).
template <bool WRITE>
__device__ __forceinline__ uint32_t mf_lane_emit_dev(const uint8_t *seg, uint32_t seglen, const uint32_t *matches, uint32_t nmatch,
                                                     uint32_t lane, const uint32_t *ll_codes, const uint8_t *ll_lens,
                                                     const uint32_t *d_codes, MfBitWriterDev *w) {
    const uint32_t lo = lane * 32u;
    const uint32_t hi = min(lo + 32u, seglen);
    if (lo >= hi) return 0;
    uint32_t wd[8];
    {
        const uint4 v0 = *reinterpret_cast<const uint4 *>(seg + lo), v1 = *reinterpret_cast<const uint4 *>(seg + lo + 16);
        wd[0] = v0.x; wd[1] = v0.y; wd[2] = v0.z; wd[3] = v0.w; wd[4] = v1.x; wd[5] = v1.y; wd[6] = v1.z; wd[7] = v1.w;
    }
    uint32_t a = 0, b = nmatch;
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (mf_match_pos(matches[mid]) < lo) a = mid + 1; else b = mid; }
    uint32_t mi = a;
    uint32_t p = lo;
    if (mi > 0) { const uint32_t pm = matches[mi - 1]; const uint32_t e = mf_match_pos(pm) + mf_match_len(pm); if (e > p) p = e; }
    uint32_t next_pos = mi < nmatch ? mf_match_pos(matches[mi]) : 0xFFFFFFFFu;
    uint32_t bits = 0;
    // literal-only span (no match starts in it, none reaches into it): no token bookkeeping at all.  Two literals
    // are joined (<= 30 bits) before they go to the bit writer, so the 64-bit accumulate runs 16 times, not 32.
    if (p == lo && next_pos >= hi && hi == lo + 32u) {
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
            const uint32_t b0 = (wd[k >> 2] >> ((k & 3) * 8)) & 0xFFu, b1 = (wd[k >> 2] >> (((k + 1) & 3) * 8)) & 0xFFu;
            if (WRITE) {
                const uint32_t e0 = ll_codes[b0], e1 = ll_codes[b1];
                const uint32_t l0 = e0 >> 16, l1 = e1 >> 16;
                w->put((e0 & 0xFFFFu) | ((e1 & 0xFFFFu) << l0), l0 + l1);
                bits += l0 + l1;
            } else {
                bits += (uint32_t)ll_lens[b0] + (uint32_t)ll_lens[b1];
            }
        }
        return bits;
    }
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const uint32_t pos = lo + (uint32_t)k;
        if (pos < hi && pos == p) {
            if (pos == next_pos) {
                const uint32_t m = matches[mi];
                const uint32_t len = mf_match_len(m), dist = mf_match_dist(m);
                uint32_t ls, lb, lv, ds, db, dv;
                mf_len_sym(len, ls, lb, lv); mf_dist_sym(dist, ds, db, dv);
                const uint32_t le = ll_codes[ls], de = d_codes[ds];
                bits += (le >> 16) + lb + (de >> 16) + db;
                if (WRITE) {
                    w->put((le & 0xFFFFu) | (lv << (le >> 16)), (le >> 16) + lb);
                    w->put((de & 0xFFFFu) | (dv << (de >> 16)), (de >> 16) + db);
                }
                p += len; mi++;
                next_pos = mi < nmatch ? mf_match_pos(matches[mi]) : 0xFFFFFFFFu;
            } else {
                const uint32_t byte = (wd[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
                if (WRITE) { const uint32_t e = ll_codes[byte]; bits += e >> 16; w->put(e & 0xFFFFu, e >> 16); }
                else bits += ll_lens[byte];
                p++;
            }
        }
    }
    return bits;
}

#define MF_ENC_THREADS 512
#define MF_ENC_WARPS 16
#define MF_HASH_BITS 10
#define MF_ENC_GROUP 64u         // consecutive chunks of a member handed to one CTA: codes may be carried inside a group
#define MF_HDR_CACHE_WORDS 192u
#define MF_REGION_B 34816u      // hash tables in pass 2 (32 KiB), packed output words afterwards

struct __align__(16) MfEncSmem {
    uint8_t data[MF_CHUNK + 32];
    uint8_t regionB[MF_REGION_B];
    uint32_t match[MF_SEGS_PER_CHUNK][MF_MAXM];
    uint32_t crc_tab[4][256];
    uint32_t ll_hist[288];
    uint32_t d_hist[32];
    uint32_t ll_codes[288];
    uint32_t d_codes[32];
    uint8_t ll_lens[288];
    uint8_t d_lens[32];
    MfWarpHdrScratch whdr;           // warp 0: litlen code, then the block header
    MfWarpCodeScratch wdist;         // warp 1: distance code
    uint32_t hdr_bits;
    uint32_t ticket;
    uint32_t nmatch[MF_SEGS_PER_CHUNK];
    uint32_t seg_bits[MF_SEGS_PER_CHUNK];
    uint32_t seg_off[MF_SEGS_PER_CHUNK];
    uint16_t lane_bits[MF_SEGS_PER_CHUNK][32];
    uint32_t crc_part[MF_ENC_WARPS];
    uint32_t total_bits;
    uint32_t comp_bytes;
    uint32_t stored;
    // the codes of the previous chunk of this group, kept while they stay within 1/128 of what a rebuild could give
    uint32_t hdr_cache[MF_HDR_CACHE_WORDS];   // its block header, bit for bit (every chunk starts byte-aligned)
    uint32_t cache_hdr_bits;                  // 0: nothing cached
    float cache_red16;                        // redundancy of the cached code on the chunk it was built for, bits*16 per symbol
    uint32_t dec_cost, dec_miss, dec_a_ll, dec_b_ll, dec_a_d, dec_b_d;   // this chunk against the cached code lengths
};

// flags bit 0: MFCAS_FLAG_NO_LZ
__global__ void __launch_bounds__(MF_ENC_THREADS, 2)
mf_deflate_kernel(const MfEncMember *__restrict__ members, const uint32_t *__restrict__ chunk_member, uint32_t n_chunks_total,
                  uint8_t *__restrict__ slots, uint32_t *__restrict__ comp_bytes_out, uint16_t *__restrict__ hdr_bits_out,
                  uint16_t *__restrict__ seg_bits_out, uint32_t *__restrict__ chunk_crc_out, uint32_t flags,
                  unsigned long long *__restrict__ phase_clk, uint32_t *__restrict__ ticket,
                  const uint32_t *__restrict__ grp_first, uint32_t n_groups) {
    extern __shared__ __align__(16) uint8_t mf_smem_raw[];
    MfEncSmem &S = *reinterpret_cast<MfEncSmem *>(mf_smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // persistent CTAs: the grid is capped below 2 CTAs x 148 SMs (host side) so that some SMs always keep
    // a free slot for the latency-critical SHA-1 CTAs of this and the following waves
    // Chunks are handed out by an atomic ticket, not by a static stride: CTAs that share their SM with SHA-1
    // CTAs run slower, and a static partition would make everyone wait for them.
    // The unit of work is a GROUP of up to MF_ENC_GROUP consecutive chunks of one member.  Building the two Huffman
    // codes and the block header is warp-serial work (half of a chunk's cycles); inside a group the codes of the
    // previous chunk are kept whenever they cost no more than 1/128 above what fresh codes could (pass 3a), which on
    // stationary data (float arrays) is nearly always.  Groups are member-relative, so a member's bytes do not
    // depend on the batch it was packed in or on which CTA took which ticket.
    (void)n_chunks_total;
    for (;;) {
        if (tid == 0) S.ticket = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t grp = S.ticket;
        if (grp >= n_groups) break;
        const uint32_t c0 = grp_first[grp];
        const MfEncMember mem = members[chunk_member[c0]];
        const uint32_t c1 = min(c0 + MF_ENC_GROUP, mem.first_chunk + mem.n_chunks);
        if (tid == 0) S.cache_hdr_bits = 0;
        uint32_t dry = 0;                        // consecutive probed windows of this warp that produced no match
    for (uint32_t c = c0; c < c1; c++) {
    // optional phase profile (MFCAS_PHASE_TIMING=1): clock64 at every phase boundary of thread 0
#define MF_PHASE(k) do { if (phase_clk && tid == 0) phase_clk[(size_t)c * 16 + (k)] = clock64(); } while (0)
    MF_PHASE(0);
    const uint32_t ci = c - mem.first_chunk;
    const uint32_t len = min((uint32_t)MF_CHUNK, mem.raw_len - ci * MF_CHUNK);
    const uint8_t *src = mem.src + (size_t)ci * MF_CHUNK;

    // ---- pass 0: stage the chunk, zero histograms, copy CRC tables
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(src);
        uint4 *s = reinterpret_cast<uint4 *>(S.data);
        uint32_t nv = (len + 15u) >> 4;
        for (uint32_t i = tid; i < nv; i += MF_ENC_THREADS) s[i] = __ldg(g + i);
        for (uint32_t i = nv + tid; i < (MF_CHUNK + 32) / 16; i += MF_ENC_THREADS) s[i] = make_uint4(0, 0, 0, 0);
        for (uint32_t i = tid; i < 1024; i += MF_ENC_THREADS) (&S.crc_tab[0][0])[i] = (&g_mf_crc_tab[0][0])[i];
        for (uint32_t i = tid; i < 288; i += MF_ENC_THREADS) S.ll_hist[i] = 0;
        if (tid < 32) S.d_hist[tid] = 0;
        if (tid == 0) { S.dec_cost = 0; S.dec_miss = 0; S.dec_a_ll = 0; S.dec_b_ll = 0; S.dec_a_d = 0; S.dec_b_d = 0; }
    }
    __syncthreads();
    // bytes past `len` inside the last 16-byte vector came from the (padded) source: clear them so
    // that nothing downstream depends on them
    if (tid < 16 && (len & 15u) && tid >= (len & 15u)) S.data[(len & ~15u) + tid] = 0;
    __syncthreads();

    MF_PHASE(1);
    MF_PHASE(2);
    // ---- pass 2: LZ77 over this warp's two segments
    {
        uint16_t *ht = reinterpret_cast<uint16_t *>(S.regionB) + warp * (1u << MF_HASH_BITS);
        const bool lz = (flags & 1u) == 0;
        for (uint32_t si = 0; si < 2; si++) {
            const uint32_t seg = warp * 2 + si;
            const uint32_t base = seg * MF_SEG;
            if (base >= len) { if (lane == 0) S.nmatch[seg] = 0; continue; }
            const uint32_t seglen = min((uint32_t)MF_SEG, len - base);
            for (uint32_t i = lane; i < (1u << MF_HASH_BITS) / 2; i += 32) reinterpret_cast<uint32_t *>(ht)[i] = 0;
            __syncwarp();
            uint32_t nm = 0;
            uint32_t *ml = S.match[seg];
            uint32_t pos = 0;
            while (pos < seglen) {
                const uint32_t p = pos + lane;
                const bool valid = p < seglen;
                // incompressible stretches (float noise): the longer nothing matched, the fewer windows are probed --
                // every 4th after 3 dry windows, every 8th after 6, every 16th after 12 (LZ4-style acceleration);
                // the rest only feed the literal histogram.  `dry` is carried across this warp's segments and
                // across the chunks of the group: a probed window costs ~80 instructions, a skipped one ~4.
                const uint32_t stride = dry >= 12u ? 512u : dry >= 6u ? 256u : 128u;
                if (!lz || (dry >= 3u && (pos & (stride - 32u)) != 0)) {
                    // all literals up to the next probed window, in one step
                    const uint32_t stop = lz ? min(seglen, (pos + stride) & ~(stride - 1u)) : seglen;
                    for (uint32_t q = p; q < stop; q += 32) atomicAdd(&S.ll_hist[S.data[base + q]], 1u);
                    pos = stop;
                    continue;
                }
                const uint32_t nm_before = nm;
                const uint32_t w = valid ? mf_ld32u(S.data, base + p) : 0u;
                const bool can = lz && (p + MF_MIN_MATCH <= seglen) && (nm < MF_MAXM);
                const uint32_t h = (w * 2654435761u) >> (32 - MF_HASH_BITS);
                int q = -1;
                if (can) { uint32_t cnd = ht[h]; if (cnd) q = (int)cnd - 1; }
                __syncwarp();
                if (can) ht[h] = (uint16_t)(p + 1);      // lanes with equal hashes write the same slot: whichever lands is a valid candidate (racecheck flags it)
                __syncwarp();
                // in-window probes (the hash table only knows earlier windows)
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
                if (can && q >= 0) has = (mf_ld32u(S.data, base + (uint32_t)q) == w);
                uint32_t bal = __ballot_sync(0xFFFFFFFFu, has);
                const uint32_t validmask = __ballot_sync(0xFFFFFFFFu, valid);
                uint32_t litmask = 0;
                uint32_t cur = 0, adv = 32;
                while (true) {
                    uint32_t m = cur < 32 ? (bal & (0xFFFFFFFFu << cur)) : 0u;
                    if (!m) { if (cur < 32) litmask |= (0xFFFFFFFFu << cur); adv = 32; break; }
                    const uint32_t L = __ffs(m) - 1;
                    const uint32_t pL = pos + L;
                    const uint32_t qL = (uint32_t)__shfl_sync(0xFFFFFFFFu, q, L);
                    const uint32_t maxlen = min((uint32_t)MF_MAX_MATCH, seglen - pL);
                    uint32_t mlen = maxlen;
                    for (uint32_t off = 4; off < maxlen; off += 32) {
                        uint32_t o = off + lane;
                        bool ne = (o >= maxlen) || (S.data[base + pL + o] != S.data[base + qL + o]);
                        uint32_t bm = __ballot_sync(0xFFFFFFFFu, ne);
                        if (bm) { mlen = off + (__ffs(bm) - 1); break; }
                    }
                    const uint32_t dist = pL - qL;
                    const bool accept = (mlen >= 6) || (mlen >= 4 && dist <= 32) || (mlen == 5 && dist <= 256);
                    if (accept) {
                        if (L > cur) litmask |= ((1u << L) - 1u) & (0xFFFFFFFFu << cur);
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
                        litmask |= ((L == 31 ? 0xFFFFFFFFu : ((1u << (L + 1)) - 1u))) & (0xFFFFFFFFu << cur);
                        cur = L + 1;
                    }
                }
                litmask &= validmask;
                if ((litmask >> lane) & 1u) atomicAdd(&S.ll_hist[w & 0xFFu], 1u);
                dry = (nm == nm_before) ? dry + 1 : 0;
                pos += adv;
            }
            if (lane == 0) S.nmatch[seg] = nm;
        }
    }
    if (tid == 0) S.ll_hist[256] = 1;       // end-of-block
    __syncthreads();

    MF_PHASE(3);
    // ---- pass 3a: price this chunk's symbols with the code lengths already in shared memory (previous chunk of the
    // group) and against the entropy bound of its own histogram.  Integer sums only: the verdict must not depend
    // on the order of the atomics.  Symbols 0..285 on warps 0..8, distance symbols on warp 9.
    if (tid < 320) {
        uint32_t h = 0, l = 0;
        if (warp < 9) { if (tid < 286) { h = S.ll_hist[tid]; l = S.ll_lens[tid]; } }
        else if (tid - 288u < 30u) { h = S.d_hist[tid - 288u]; l = S.d_lens[tid - 288u]; }
        const uint32_t miss = __ballot_sync(0xFFFFFFFFu, h != 0 && l == 0);
        const uint32_t cost = mf_warp_sum(h * l), a = mf_warp_sum(h);
        const uint32_t bb = mf_warp_sum(h > 1 ? (uint32_t)((float)h * __log2f((float)h) * 16.f + 0.5f) : 0u);
        if (lane == 0) {
            atomicAdd(&S.dec_cost, cost);
            if (miss) atomicOr(&S.dec_miss, 1u);
            if (warp < 9) { atomicAdd(&S.dec_a_ll, a); atomicAdd(&S.dec_b_ll, bb); }
            else { atomicAdd(&S.dec_a_d, a); atomicAdd(&S.dec_b_d, bb); }
        }
    }
    __syncthreads();
    // entropy bound of the chunk's two histograms, in bits * 16 (every thread computes the same value)
    float e16 = 0.f;
    { const uint32_t a = S.dec_a_ll, bsum = S.dec_b_ll; if (a > 1) e16 += (float)a * __log2f((float)a) * 16.f - (float)bsum; }
    { const uint32_t a = S.dec_a_d, bsum = S.dec_b_d; if (a > 1) e16 += (float)a * __log2f((float)a) * 16.f - (float)bsum; }
    e16 = fmaxf(e16, 0.f);
    const float nsym = (float)(S.dec_a_ll + S.dec_a_d);
    const uint32_t cached_hb = S.cache_hdr_bits;
    const bool reuse = cached_hb != 0 && S.dec_miss == 0 &&
                       (float)S.dec_cost * 16.f <= (e16 + S.cache_red16 * nsym) * (1.f + 1.f / 128.f) + 1024.f;
    uint32_t *outw = reinterpret_cast<uint32_t *>(S.regionB);
    if (reuse) {
        // ---- pass 3 (codes kept): all warps zero region B (dead as hash tables now) and take one 64-byte CRC piece
        // per thread; then the cached header is copied and the bits are counted with the cached code lengths
        uint4 *z = reinterpret_cast<uint4 *>(S.regionB);
        for (uint32_t i = tid; i < MF_REGION_B / 16; i += MF_ENC_THREADS) z[i] = make_uint4(0, 0, 0, 0);
        uint32_t acc = 0;
        {
            const uint32_t start = tid * 64u;
            if (start < len) {
                const uint32_t end = min(start + 64u, len);
                const uint32_t nw = (end - start) & ~3u;
                uint32_t crc = mf_crc_words(S.data, start, nw, S.crc_tab);
                for (uint32_t p = start + nw; p < end; p++) crc = S.crc_tab[0][(crc ^ S.data[p]) & 0xFFu] ^ (crc >> 8);
                const uint32_t mul = (len == MF_CHUNK) ? g_mf_piece_mul[511u - tid] : mf_xpow8(len - end);
                acc = mf_gf2_mul(crc, mul);
            }
        }
        acc = mf_warp_xor(acc);
        if (lane == 0) S.crc_part[warp] = acc;
        MF_PHASE(8);
        __syncthreads();
        MF_PHASE(9); MF_PHASE(14); MF_PHASE(15); MF_PHASE(10); MF_PHASE(11); MF_PHASE(12);
        if (warp == 0) {
            for (uint32_t i = lane; i < ((cached_hb + 31u) >> 5); i += 32) outw[i] = S.hdr_cache[i];
            if (lane == 0) S.hdr_bits = cached_hb;
        }
        for (uint32_t si = 0; si < 2; si++) {
            const uint32_t seg = warp * 2 + si;
            const uint32_t base = seg * MF_SEG;
            uint32_t bits = 0;
            if (base < len) {
                const uint32_t seglen = min((uint32_t)MF_SEG, len - base);
                bits = mf_lane_emit_dev<false>(S.data + base, seglen, S.match[seg], S.nmatch[seg], lane, S.ll_codes, S.ll_lens, S.d_codes, nullptr);
            }
            S.lane_bits[seg][lane] = (uint16_t)bits;
            uint32_t incl = mf_warp_incl_scan(bits, lane);
            if (lane == 31) S.seg_bits[seg] = incl;
        }
    } else {
    // ---- pass 3 (fresh codes): Huffman codes, block header
    // warp 0: literal/length code; warp 1: distance code (warp-synchronous, in registers); warps 2..15 meanwhile
    // zero region B (dead as hash tables now) for the bit writers and compute the chunk's CRC-32
    // (64-byte piece per lane, slice-by-4 tables in smem, GF(2) shift to the piece's position, xor-reduce).
    if (warp == 0) { mf_build_code_warp<9>(S.ll_hist, 286, 15, S.ll_lens, S.ll_codes, S.whdr.code, lane); if (lane == 0) S.crc_part[0] = 0; }
    else if (warp == 1) { mf_build_code_warp<1>(S.d_hist, 30, 15, S.d_lens, S.d_codes, S.wdist, lane); if (lane == 0) S.crc_part[1] = 0; }
    else {
        uint4 *z = reinterpret_cast<uint4 *>(S.regionB);
        for (uint32_t i = tid - 64; i < MF_REGION_B / 16; i += MF_ENC_THREADS - 64) z[i] = make_uint4(0, 0, 0, 0);
        uint32_t acc = 0;
        for (uint32_t g = tid - 64; g < 512u; g += MF_ENC_THREADS - 64) {
            const uint32_t start = g * 64u;
            if (start < len) {
                const uint32_t end = min(start + 64u, len);
                const uint32_t nw = (end - start) & ~3u;
                uint32_t crc = mf_crc_words(S.data, start, nw, S.crc_tab);
                for (uint32_t p = start + nw; p < end; p++) crc = S.crc_tab[0][(crc ^ S.data[p]) & 0xFFu] ^ (crc >> 8);
                const uint32_t mul = (len == MF_CHUNK) ? g_mf_piece_mul[511u - g] : mf_xpow8(len - end);
                acc ^= mf_gf2_mul(crc, mul);
            }
        }
        acc = mf_warp_xor(acc);
        if (lane == 0) S.crc_part[warp] = acc;
    }
    MF_PHASE(8);
    __syncthreads();
    MF_PHASE(9);
    // warp 0: the block header (needs both codes); warps 1..15: bit counts per lane and per segment (need only
    // the code lengths), so the header's serial latency hides behind the counting
    if (warp == 0) {
        uint32_t hb = mf_header_warp(S.ll_lens, S.d_lens, outw, S.whdr, lane, phase_clk ? &phase_clk[(size_t)c * 16 + 10] : nullptr);
        if (lane == 0) S.hdr_bits = hb;
        // keep the header and the redundancy of these codes on their own chunk for the next chunk of the group
        uint32_t own = 0;
        for (uint32_t k = lane; k < 286u; k += 32) own += S.ll_hist[k] * S.ll_lens[k];
        if (lane < 30) own += S.d_hist[lane] * S.d_lens[lane];
        own = mf_warp_sum(own);
        const bool fits = hb <= MF_HDR_CACHE_WORDS * 32u;
        if (fits) for (uint32_t i = lane; i < ((hb + 31u) >> 5); i += 32) S.hdr_cache[i] = outw[i];
        if (lane == 0) {
            S.cache_hdr_bits = fits ? hb : 0u;
            S.cache_red16 = nsym > 0.f ? fmaxf((float)own * 16.f - e16, 0.f) / nsym : 0.f;
        }
    } else {
        for (uint32_t seg = warp - 1; seg < MF_SEGS_PER_CHUNK; seg += MF_ENC_WARPS - 1) {
            const uint32_t base = seg * MF_SEG;
            uint32_t bits = 0;
            if (base < len) {
                const uint32_t seglen = min((uint32_t)MF_SEG, len - base);
                bits = mf_lane_emit_dev<false>(S.data + base, seglen, S.match[seg], S.nmatch[seg], lane, S.ll_codes, S.ll_lens, S.d_codes, nullptr);
            }
            S.lane_bits[seg][lane] = (uint16_t)bits;
            uint32_t incl = mf_warp_incl_scan(bits, lane);
            if (lane == 31) S.seg_bits[seg] = incl;
        }
    }
    }
    MF_PHASE(13);
    __syncthreads();
    MF_PHASE(4);
    if (warp == 0) {
        uint32_t b = S.seg_bits[lane];
        uint32_t incl = mf_warp_incl_scan(b, lane);
        S.seg_off[lane] = S.hdr_bits + incl - b;
        if (lane == 31) {
            uint32_t eob = S.ll_codes[256];
            uint32_t end_tokens = S.hdr_bits + incl;            // where EOB goes
            uint32_t end_bits = end_tokens + (eob >> 16) + 3u;       // + empty stored block header (000)
            uint32_t dyn_bytes = ((end_bits + 7u) >> 3) + 4u;        // + LEN/NLEN of the sync marker
            S.total_bits = end_tokens;
            if (dyn_bytes >= len + 5u) { S.stored = 1; S.comp_bytes = len + 5u; }
            else { S.stored = 0; S.comp_bytes = dyn_bytes; }
        }
    }
    __syncthreads();

    MF_PHASE(5);
    // ---- pass 5: pack
    if (!S.stored) {
        for (uint32_t si = 0; si < 2; si++) {
            const uint32_t seg = warp * 2 + si;
            const uint32_t base = seg * MF_SEG;
            if (base >= len) continue;
            const uint32_t seglen = min((uint32_t)MF_SEG, len - base);
            uint32_t bits = S.lane_bits[seg][lane];
            uint32_t incl = mf_warp_incl_scan(bits, lane);
            if (bits) {
                MfBitWriterDev bw;
                bw.init(outw, S.seg_off[seg] + incl - bits);
                mf_lane_emit_dev<true>(S.data + base, seglen, S.match[seg], S.nmatch[seg], lane, S.ll_codes, S.ll_lens, S.d_codes, &bw);
                bw.flush();
            }
        }
        if (tid == 0) {
            MfBitWriter bw;
            bw.init(outw, S.total_bits);
            uint32_t eob = S.ll_codes[256];
            bw.put(eob & 0xFFFFu, eob >> 16);
            bw.flush();
            // sync marker: 3 zero bits, pad to byte, 00 00 FF FF (words are zero already)
            uint32_t end_bits = S.total_bits + (eob >> 16) + 3u;
            uint32_t byte = ((end_bits + 7u) >> 3) + 2u;
            uint8_t *ob = S.regionB;
            atomicOr(&outw[byte >> 2], 0xFFu << ((byte & 3u) * 8u));
            atomicOr(&outw[(byte + 1) >> 2], 0xFFu << (((byte + 1) & 3u) * 8u));
            (void)ob;
        }
    } else {
        // stored block: 00 | LEN | NLEN | raw bytes (byte-aligned by construction)
        uint32_t nwords = (len + 5u + 3u) >> 2;
        for (uint32_t j = tid; j < nwords; j += MF_ENC_THREADS) {
            uint32_t v;
            if (j == 0) v = 0x00u | ((len & 0xFFFFu) << 8) | (((~len) & 0xFFu) << 24);
            else if (j == 1) v = (((~len) >> 8) & 0xFFu) | ((uint32_t)S.data[0] << 8) | ((uint32_t)S.data[1] << 16) | ((uint32_t)S.data[2] << 24);
            else v = mf_ld32u(S.data, 4u * j - 5u);
            outw[j] = v;
        }
    }
    __syncthreads();

    MF_PHASE(6);
    // ---- write out: slot bytes, index entries, chunk CRC
    {
        const uint32_t cb = S.comp_bytes;
        uint4 *g = reinterpret_cast<uint4 *>(slots + (size_t)c * MF_SLOT);
        const uint4 *s = reinterpret_cast<const uint4 *>(S.regionB);
        for (uint32_t i = tid; i < ((cb + 15u) >> 4); i += MF_ENC_THREADS) g[i] = s[i];
        if (tid < MF_SEGS_PER_CHUNK) seg_bits_out[(size_t)c * MF_SEGS_PER_CHUNK + tid] = S.stored ? 0 : (uint16_t)S.seg_bits[tid];
        if (tid == 0) {
            comp_bytes_out[c] = cb;
            hdr_bits_out[c] = S.stored ? (uint16_t)MF_STORED_MARK : (uint16_t)S.hdr_bits;
            uint32_t crc = 0;
            for (int k = 0; k < MF_ENC_WARPS; k++) crc ^= S.crc_part[k];
            chunk_crc_out[c] = crc;
        }
    }
    MF_PHASE(7);
    __syncthreads();          // shared memory is reused by the next chunk
    }
    }
#undef MF_PHASE
}

// ------------------------------------------------------------------ gather: members -> final bytes
__device__ __forceinline__ void mf_put_le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
__device__ __forceinline__ void mf_put_le32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

#define MF_GA_THREADS 256
// One CTA per blob: lays out its members, writes gzip headers + index + trailers, records where
// every chunk's bytes go.  chunk_dst[c] = byte offset inside the blob's dst, or ~0 when the blob
// does not fit dst_cap (status DST_TOO_SMALL).
__global__ void __launch_bounds__(MF_GA_THREADS)
mf_gather_layout_kernel(const MfEncBlob *__restrict__ blobs, const MfEncMember *__restrict__ members,
                        const uint32_t *__restrict__ comp_bytes, const uint16_t *__restrict__ hdr_bits,
                        const uint16_t *__restrict__ seg_bits, const uint32_t *__restrict__ chunk_crc,
                        uint64_t *__restrict__ chunk_dst, uint64_t *__restrict__ dst_len, int32_t *__restrict__ status) {
    __shared__ uint32_t s_scan[MF_GA_THREADS];
    __shared__ uint32_t s_crc[MF_GA_THREADS];
    __shared__ uint64_t s_off;
    __shared__ int s_fit;
    const uint32_t b = blockIdx.x, tid = threadIdx.x;
    const MfEncBlob blob = blobs[b];
    // first walk: total size (so that an undersized dst is never written)
    if (tid == 0) { s_off = 0; }
    __syncthreads();
    uint64_t total = 0;
    for (uint32_t mi = 0; mi < blob.n_members; mi++) {
        const MfEncMember mem = members[blob.first_member + mi];
        uint32_t part = 0;
        for (uint32_t k = tid; k < mem.n_chunks; k += MF_GA_THREADS) part += comp_bytes[mem.first_chunk + k];
        s_scan[tid] = part;
        __syncthreads();
        for (int d = MF_GA_THREADS / 2; d > 0; d >>= 1) { if ((int)tid < d) s_scan[tid] += s_scan[tid + d]; __syncthreads(); }
        total += (uint64_t)mf_member_header_bytes(mem.n_chunks) + s_scan[0] + MF_TRAILER;
        __syncthreads();
    }
    if (tid == 0) {
        s_fit = total <= blob.dst_cap;
        dst_len[b] = total;
        status[b] = s_fit ? 0 : 1 /* MFCAS_E_DST_TOO_SMALL */;
    }
    __syncthreads();
    const bool fit = s_fit != 0;
    uint64_t off = 0;
    for (uint32_t mi = 0; mi < blob.n_members; mi++) {
        const MfEncMember mem = members[blob.first_member + mi];
        const uint32_t nc = mem.n_chunks;
        const uint32_t hdr = mf_member_header_bytes(nc);
        // exclusive scan of comp_bytes over this member's chunks (<= 768 = 3 per thread, blocked)
        const uint32_t per = (nc + MF_GA_THREADS - 1) / MF_GA_THREADS;
        uint32_t mine = 0;
        for (uint32_t k = 0; k < per; k++) { uint32_t ch = tid * per + k; if (ch < nc) mine += comp_bytes[mem.first_chunk + ch]; }
        s_scan[tid] = mine;
        __syncthreads();
        for (int d = 1; d < MF_GA_THREADS; d <<= 1) {
            uint32_t v = (int)tid >= d ? s_scan[tid - d] : 0u;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const uint32_t payload = s_scan[MF_GA_THREADS - 1];
        uint32_t run = s_scan[tid] - mine;                 // exclusive prefix of this thread's first chunk
        uint32_t crc_acc = 0;
        uint8_t *mbase = blob.dst + off;
        for (uint32_t k = 0; k < per; k++) {
            uint32_t ch = tid * per + k;
            if (ch >= nc) break;
            const uint32_t gc = mem.first_chunk + ch;
            const uint32_t cb = comp_bytes[gc];
            chunk_dst[gc] = fit ? (off + hdr + run) : ~0ull;
            // CRC contribution: chunk raw CRC shifted left by the raw bytes that follow it
            uint64_t chunk_end = min((uint64_t)(ch + 1) * MF_CHUNK, (uint64_t)mem.raw_len);
            crc_acc ^= mf_gf2_mul(chunk_crc[gc], mf_xpow8((uint64_t)mem.raw_len - chunk_end));
            if (fit) {
                uint8_t *e = mbase + MF_HDR_FIXED + (size_t)ch * 8u;
                mf_put_le32(e, cb); mf_put_le16(e + 4, hdr_bits[gc]); mf_put_le16(e + 6, 0);
                uint8_t *sb = mbase + MF_HDR_FIXED + (size_t)nc * 8u + (size_t)ch * 64u;
                for (uint32_t s = 0; s < MF_SEGS_PER_CHUNK; s++) mf_put_le16(sb + 2 * s, seg_bits[(size_t)gc * MF_SEGS_PER_CHUNK + s]);
            }
            run += cb;
        }
        s_crc[tid] = crc_acc;
        __syncthreads();
        for (int d = MF_GA_THREADS / 2; d > 0; d >>= 1) { if ((int)tid < d) s_crc[tid] ^= s_crc[tid + d]; __syncthreads(); }
        if (tid == 0 && fit) {
            uint8_t *h = mbase;
            h[0] = 0x1f; h[1] = 0x8b; h[2] = 8; h[3] = 4 /* FEXTRA */;
            h[4] = h[5] = h[6] = h[7] = 0;              /* MTIME 0: deterministic output */
            h[8] = 0; h[9] = 255;                        /* XFL 0, OS unknown (as gzip.py writes) */
            mf_put_le16(h + 10, hdr - 12);               /* XLEN */
            h[12] = 'M'; h[13] = 'F'; mf_put_le16(h + 14, hdr - 16);
            h[16] = MF_FORMAT_VERSION; h[17] = MF_CHUNK_LOG2; h[18] = MF_SEG_LOG2; h[19] = 0;
            mf_put_le32(h + 20, mem.raw_len); mf_put_le32(h + 24, nc); mf_put_le32(h + 28, 0);
            uint8_t *t = mbase + hdr + payload;
            t[0] = 1; t[1] = 0; t[2] = 0; t[3] = 0xFF; t[4] = 0xFF;   /* final empty stored block */
            mf_put_le32(t + 5, mf_crc_finish(s_crc[0], mem.raw_len));
            mf_put_le32(t + 9, mem.raw_len);
        }
        off += (uint64_t)hdr + payload + MF_TRAILER;
        __syncthreads();
    }
}

// One CTA per chunk: slot (16-byte aligned) -> final position (any alignment).
__global__ void __launch_bounds__(256)
mf_gather_copy_kernel(const MfEncBlob *__restrict__ blobs, const MfEncMember *__restrict__ members,
                      const uint32_t *__restrict__ chunk_member, uint32_t n_chunks_total,
                      const uint8_t *__restrict__ slots, const uint32_t *__restrict__ comp_bytes,
                      const uint64_t *__restrict__ chunk_dst) {
    const uint32_t c = blockIdx.x;
    if (c >= n_chunks_total) return;
    const uint64_t doff = chunk_dst[c];
    if (doff == ~0ull) return;
    uint8_t *dst = blobs[members[chunk_member[c]].blob].dst + doff;
    const uint8_t *src = slots + (size_t)c * MF_SLOT;
    const uint32_t n = comp_bytes[c];
    const uint32_t tid = threadIdx.x;
    // head: bytes until dst is 16-byte aligned
    uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
    if (head > n) head = n;
    if (tid < head) dst[tid] = src[tid];
    const uint32_t body = (n - head) >> 4;           // whole 16-byte vectors
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src);
    uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
    const uint32_t sh = (head & 3u) * 8u;
    const uint32_t w0 = head >> 2;
    for (uint32_t v = tid; v < body; v += blockDim.x) {
        const uint32_t i = w0 + 4u * v;
        uint32_t a = sw[i], b = sw[i + 1], cc = sw[i + 2], d = sw[i + 3], e = sh ? sw[i + 4] : 0u;
        uint4 o;
        o.x = __funnelshift_r(a, b, sh); o.y = __funnelshift_r(b, cc, sh);
        o.z = __funnelshift_r(cc, d, sh); o.w = __funnelshift_r(d, e, sh);
        dv[v] = o;
    }
    const uint32_t done = head + (body << 4);
    if (tid < n - done) dst[done + tid] = src[done + tid];
}
