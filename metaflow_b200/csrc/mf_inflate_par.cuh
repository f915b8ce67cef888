// mf_inflate_par.cuh -- the CUDA kernels around mf_inflate_par.h: block-parallel decoding of foreign gzip streams
// (files written by stock Metaflow, content_addressed_store.py:211-216, or any other RFC 1952 writer).
//
//   mf_fscan_kernel      thread = 512 bytes of compressed input: every bit position tested for a dynamic block header
//   mf_fcompact_kernel   CTA per blob: the hits of its ranges -> one sorted candidate list
//   mf_fsize_kernel      lane per candidate: where the block ends and how much it produces (tables in shared memory)
//   mf_fwalk_kernel      warp per blob: the true chain of blocks, their output offsets, the members
//   mf_fdecode_kernel    lane per block: 16-bit symbols (bytes and window markers) at the block's output offset
//   mf_fresolve_kernel   CTA per blob, blocks in order: markers -> bytes, into the destination
//   mf_fverify_kernel    CTA per blob: CRC-32 / ISIZE of every member
// A blob whose chain cannot be followed (more candidates than slots, a block no candidate matches, ...) keeps kind 1
// and is decoded by the serial kernel (mf_gunzip_lane_kernel) as before: slower, never wrong.
#pragma once
#include <cuda_runtime.h>
#include "mf_inflate_par.h"
#include "mf_inflate.cuh"

#define MF_F_RANGE_BYTES 512u
#define MF_F_RANGE_SLOTS 2u              // candidates kept per range (more -> the blob falls back to the serial decoder)
#define MF_F_EXTRA_BLOCKS 64u            // block slots per blob beyond one per range (stored / fixed blocks, tiny streams)
#define MF_F_MAX_MEMBERS 64u
#define MF_F_PRE 24u                     // positions a scan thread collects before it validates them
#define MF_F_LANES 96u                   // decoder lanes per CTA: 2 KiB of tables each

struct MfFBlob {
    const uint8_t *src; uint64_t src_len;
    uint8_t *dst; uint64_t dst_cap;
    uint32_t blob;            // index in the wave (kind / status / dst_len)
    uint32_t range_base;      // first scan range == first candidate slot
    uint32_t n_ranges;
    uint32_t blk_base;        // first block slot (capacity n_ranges + MF_F_EXTRA_BLOCKS)
    uint64_t out16_off;       // first symbol of its region of the 16-bit scratch (capacity dst_cap symbols)
};
struct MfSpecLane { MfSpecWork w; uint32_t pad; };   // +4 bytes: lanes at a 2048-byte stride would share one bank (32-way conflicts)
struct MfFState { uint32_t n_cands, n_blocks, n_members, mode; uint64_t total_out; };   // mode: 0 pending, 1 parallel, 2 fall back, 3 failed

// Device forms of mf_spec_size_block / mf_spec_decode_block (mf_inflate_par.h keeps the portable ones: the CPU model
// runs those).  Same tables, same checks, same results; the symbol loop reads the bit stream a 32-bit word at a time with
// two words of lookahead (MfFastReader) instead of byte by byte -- a lane decoding alone hides no latency, so the
// length of the per-symbol dependency chain is what its speed is.
__device__ int mf_spec_symbols_dev(const uint8_t *src, uint64_t n, uint64_t first_sym_bit, const MfSpecWork &W, uint64_t cap, bool exact_cap,
                                   uint16_t *out16, uint64_t *end_bit, uint64_t *out_len) {
    const uint64_t avail = n * 8ull - first_sym_bit;
    MfFastReader fr; fr.init(src, first_sym_bit, src + n);
    uint64_t op = 0;
    for (;;) {
        fr.fill();
        const uint32_t e = W.lut.e[fr.peek(MF_GEN_LUT_BITS)];
        int sym;
        if (e) { sym = (int)(e & 511u); fr.drop(e >> 9); }
        else { sym = mf_canon_decode_fast(fr, W.ll.count, W.ll.symbol); if (sym < 0) return fr.used > avail ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM; }
        if (sym < 256) {
            if (op >= cap) return exact_cap ? MF_DEC_E_STREAM : MF_DEC_E_DST;
            if (out16) out16[op] = (uint16_t)sym;
            op++;
            continue;
        }
        if (fr.used > avail) return MF_DEC_E_TRUNC;
        if (sym == 256) break;
        sym -= 257;
        if (sym >= 29) return MF_DEC_E_STREAM;
        fr.fill();
        const uint32_t len = mf_lbase(sym) + fr.get(mf_lext(sym));
        fr.fill();
        const int ds = mf_canon_decode_fast(fr, W.dd.count, W.dd.symbol);
        if (ds < 0 || ds >= 30) return fr.used > avail ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        fr.fill();
        const uint32_t dist = mf_dbase(ds) + fr.get(mf_dext(ds));
        if (fr.used > avail) return MF_DEC_E_TRUNC;
        if (dist > op + 32768u) return MF_DEC_E_STREAM;          // reaches further back than any window allows
        if (op + len > cap) return exact_cap ? MF_DEC_E_STREAM : MF_DEC_E_DST;
        if (out16) {
            for (uint32_t k = 0; k < len; k++) {
                uint16_t v;
                if (dist <= op) v = out16[op - dist];
                else v = (uint16_t)(MF_SPEC_MARK | (uint32_t)(dist - op - 1u));     // that many bytes before the block's first byte
                out16[op++] = v;
            }
        } else op += len;
    }
    *end_bit = first_sym_bit + fr.used;
    *out_len = op;
    return 0;
}

__device__ int mf_spec_size_block_dev(const uint8_t *src, uint64_t n, uint64_t start_bit, uint64_t max_out, MfSpecWork &W, MfSpecCand *out) {
    MfBitReader br; br.init(src, n, start_bit);
    uint32_t last = 0;
    const int e = mf_spec_dynamic_tables(br, W, &last);
    if (e) return e;
    uint64_t end_bit = 0, made = 0;
    const int e2 = mf_spec_symbols_dev(src, n, br.bitpos(), W, max_out, false, nullptr, &end_bit, &made);
    if (e2) return e2;
    out->start_bit = start_bit; out->end_bit = end_bit; out->out_len = made;
    out->flags = MF_SPEC_F_VALID | (last ? MF_SPEC_F_FINAL : 0u);
    out->pad = 0;
    return 0;
}

__device__ int mf_spec_decode_block_dev(const uint8_t *src, uint64_t n, const MfSpecBlock &blk, MfSpecWork &W, uint16_t *out16) {
    MfBitReader br; br.init(src, n, blk.start_bit);
    if (blk.kind == MF_BLK_FIXED) { br.get(3); mf_spec_fixed_tables(W); }
    else { uint32_t last; const int e = mf_spec_dynamic_tables(br, W, &last); if (e) return e; }
    uint64_t end_bit = 0, made = 0;
    const int e2 = mf_spec_symbols_dev(src, n, br.bitpos(), W, blk.out_len, true, out16, &end_bit, &made);
    if (e2) return e2;
    return made == blk.out_len ? 0 : MF_DEC_E_STREAM;
}

__device__ __forceinline__ uint32_t mf_f_find(const MfFBlob *fb, uint32_t nf, uint32_t slot, bool by_block) {
    uint32_t lo = 0, hi = nf;                  // last blob whose base <= slot
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t base = by_block ? fb[mid].blk_base : fb[mid].range_base;
        if (base <= slot) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint64_t mf_f_load64(const uint8_t *p, const uint8_t *end) {
    uint64_t v = 0;
    if (p + 8 <= end && (((uintptr_t)p) & 7u) == 0) return *reinterpret_cast<const uint64_t *>(p);
    for (int k = 0; k < 8; k++) if (p + k < end) v |= (uint64_t)p[k] << (8 * k);
    return v;
}

__global__ void __launch_bounds__(128)
mf_fscan_kernel(const MfFBlob *__restrict__ fb, uint32_t nf, uint32_t total_ranges, uint32_t *__restrict__ range_cnt, uint64_t *__restrict__ range_pos) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    // (no early return: the warp-wide votes below need every lane)
    const bool in_grid = r < total_ranges;
    const MfFBlob B = fb[mf_f_find(fb, nf, in_grid ? r : 0u, false)];
    const uint32_t rl = (in_grid ? r : 0u) - B.range_base;
    const bool active = in_grid && rl < B.n_ranges;
    const uint8_t *src = B.src, *end = B.src + B.src_len;
    const uint64_t b0 = (uint64_t)rl * MF_F_RANGE_BYTES;
    const uint64_t b1 = active ? min(b0 + MF_F_RANGE_BYTES, B.src_len) : b0;
    uint32_t cnt = 0;
    // Two phases, so that the expensive check runs converged: first collect the positions that pass the cheap tests
    // (about 1 in 500) into a small list, then all lanes validate their k-th entry together.  (Validating on the spot
    // made 31 lanes wait whenever one lane had a hit.)
    uint16_t pre[MF_F_PRE];
    uint64_t byte = b0;
    uint64_t lo = active ? mf_f_load64(src + b0, end) : 0ull, mid = active ? mf_f_load64(src + b0 + 8, end) : 0ull, hi = 0;
    uint32_t k0 = 0;
    bool more = active && byte < b1;
    while (__any_sync(0xFFFFFFFFu, more)) {
        uint32_t np = 0;
        while (more && np < MF_F_PRE) {
            if (k0 == 0) hi = mf_f_load64(src + byte + 16, end);
            uint32_t k = k0;
            for (; k < 64 && np < MF_F_PRE; k++) {
                if (byte * 8 + k + 17 > B.src_len * 8ull || byte + (k >> 3) >= b1) { k = 64; break; }
                const uint64_t w = k ? ((lo >> k) | (mid << (64 - k))) : lo;
                if (!mf_spec_quick((uint32_t)w & 0x1FFFFu)) continue;
                const uint64_t w2 = k ? ((mid >> k) | (hi << (64 - k))) : mid;
                const uint64_t cl = (w >> 17) | (w2 << 47);                   // 57 bits from position + 17
                if (!mf_spec_clcode_complete((((uint32_t)w >> 13) & 15u) + 4u, cl)) continue;
                pre[np++] = (uint16_t)((byte - b0) * 8 + k);
            }
            if (k >= 64) { k0 = 0; byte += 8; lo = mid; mid = hi; more = byte < b1; }
            else k0 = k;                                                      // list full: resume here after validating
        }
        uint32_t maxn = np;
        for (int d = 16; d > 0; d >>= 1) maxn = max(maxn, __shfl_xor_sync(0xFFFFFFFFu, maxn, d));
        for (uint32_t i = 0; i < maxn; i++) {
            if (i < np) {
                const uint64_t bit = b0 * 8 + pre[i];
                if (mf_spec_header_valid(src, B.src_len, bit)) {
                    if (cnt < MF_F_RANGE_SLOTS) range_pos[(size_t)r * MF_F_RANGE_SLOTS + cnt] = bit;
                    cnt++;
                }
            }
        }
    }
    if (active) range_cnt[r] = cnt;
}

__global__ void __launch_bounds__(1024)
mf_fcompact_kernel(const MfFBlob *__restrict__ fb, const uint32_t *__restrict__ range_cnt, const uint64_t *__restrict__ range_pos,
                   MfSpecCand *__restrict__ cands, MfFState *__restrict__ state) {
    __shared__ uint32_t s_scan[1024];
    __shared__ uint32_t s_base, s_over;
    const MfFBlob B = fb[blockIdx.x];
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { s_base = 0; s_over = 0; }
    __syncthreads();
    for (uint32_t r0 = 0; r0 < B.n_ranges; r0 += 1024) {
        const uint32_t r = r0 + tid;
        uint32_t c = r < B.n_ranges ? range_cnt[B.range_base + r] : 0u;
        if (c > MF_F_RANGE_SLOTS) { s_over = 1; c = MF_F_RANGE_SLOTS; }
        s_scan[tid] = c;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) { const uint32_t v = (int)tid >= d ? s_scan[tid - d] : 0u; __syncthreads(); s_scan[tid] += v; __syncthreads(); }
        const uint32_t at = s_base + s_scan[tid] - c;
        for (uint32_t k = 0; k < c; k++) {
            if (at + k < B.n_ranges) {                          // the candidate list has one slot per range
                MfSpecCand cd; cd.start_bit = range_pos[(size_t)(B.range_base + r) * MF_F_RANGE_SLOTS + k]; cd.end_bit = 0; cd.out_len = 0; cd.flags = 0; cd.pad = 0;
                cands[B.range_base + at + k] = cd;
            } else s_over = 1;
        }
        __syncthreads();
        if (tid == 0) s_base += s_scan[1023];
        __syncthreads();
    }
    if (tid == 0) {
        MfFState st; st.n_cands = min(s_base, B.n_ranges); st.n_blocks = 0; st.n_members = 0; st.mode = s_over ? 2u : 0u; st.total_out = 0;
        state[blockIdx.x] = st;
    }
}

__global__ void __launch_bounds__(MF_F_LANES)
mf_fsize_kernel(const MfFBlob *__restrict__ fb, uint32_t nf, uint32_t total_slots, MfSpecCand *__restrict__ cands, const MfFState *__restrict__ state) {
    extern __shared__ __align__(16) uint8_t mf_smem_raw[];
    MfSpecWork &W = reinterpret_cast<MfSpecLane *>(mf_smem_raw)[threadIdx.x].w;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_slots) return;
    const uint32_t f = mf_f_find(fb, nf, g, false);
    const MfFBlob B = fb[f];
    const uint32_t i = g - B.range_base;
    if (state[f].mode != 0 || i >= state[f].n_cands) return;
    MfSpecCand c;
    const uint64_t start = cands[g].start_bit;
    if (mf_spec_size_block_dev(B.src, B.src_len, start, B.dst_cap, W, &c) == 0) cands[g] = c;       // otherwise flags stay 0: not a block
}

__global__ void __launch_bounds__(32)
mf_fwalk_kernel(const MfFBlob *__restrict__ fb, const MfSpecCand *__restrict__ cands, MfFState *__restrict__ state, MfSpecBlock *__restrict__ blocks,
                MfSpecMember *__restrict__ members, uint8_t *__restrict__ kind, uint64_t *__restrict__ dst_len, int32_t *__restrict__ status) {
    __shared__ MfSpecWork W;
    if (threadIdx.x != 0) return;
    const uint32_t f = blockIdx.x;
    const MfFBlob B = fb[f];
    MfFState st = state[f];
    if (st.mode != 0) return;                                      // falls back: kind stays 1
    uint32_t nb = 0, nm = 0;
    uint64_t total = 0;
    const int e = mf_spec_walk(B.src, B.src_len, cands + B.range_base, st.n_cands, B.dst_cap, W, blocks + B.blk_base, B.n_ranges + MF_F_EXTRA_BLOCKS, &nb,
                               members + (size_t)f * MF_F_MAX_MEMBERS, MF_F_MAX_MEMBERS, &nm, &total);
    if (e == MF_SPEC_FALLBACK) { st.mode = 2; state[f] = st; return; }
    if (e != 0) {
        st.mode = 3; state[f] = st;
        kind[B.blob] = 2; status[B.blob] = e; dst_len[B.blob] = (e == MF_DEC_E_DST) ? total : 0;
        return;
    }
    st.mode = 1; st.n_blocks = nb; st.n_members = nm; st.total_out = total;
    state[f] = st;
    kind[B.blob] = 3;                                              // the serial kernel skips it
    dst_len[B.blob] = total;
}

__global__ void __launch_bounds__(MF_F_LANES)
mf_fdecode_kernel(const MfFBlob *__restrict__ fb, uint32_t nf, uint32_t total_slots, const MfSpecBlock *__restrict__ blocks, MfFState *__restrict__ state,
                  uint16_t *__restrict__ out16) {
    extern __shared__ __align__(16) uint8_t mf_smem_raw[];
    MfSpecWork &W = reinterpret_cast<MfSpecLane *>(mf_smem_raw)[threadIdx.x].w;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_slots) return;
    const uint32_t f = mf_f_find(fb, nf, g, true);
    const MfFBlob B = fb[f];
    const uint32_t i = g - B.blk_base;
    if (state[f].mode != 1 || i >= state[f].n_blocks) return;
    const MfSpecBlock blk = blocks[g];
    if (blk.kind == MF_BLK_STORED) {                                // plain bytes: straight to the destination
        const uint8_t *p = B.src + (blk.start_bit >> 3);
        uint8_t *d = B.dst + blk.out_off;
        for (uint64_t k = 0; k < blk.out_len; k++) d[k] = p[k];
        return;
    }
    if (mf_spec_decode_block_dev(B.src, B.src_len, blk, W, out16 + B.out16_off + blk.out_off) != 0) atomicExch(&state[f].mode, 4u);   // resolve reports it
}

__global__ void __launch_bounds__(1024)
mf_fresolve_kernel(const MfFBlob *__restrict__ fb, const MfSpecBlock *__restrict__ blocks, const MfSpecMember *__restrict__ members,
                   MfFState *__restrict__ state, const uint16_t *__restrict__ out16, uint8_t *__restrict__ kind, uint64_t *__restrict__ dst_len,
                   int32_t *__restrict__ status) {
    __shared__ int s_bad;
    const uint32_t f = blockIdx.x, tid = threadIdx.x;
    const MfFBlob B = fb[f];
    const uint32_t mode = state[f].mode;
    if (mode != 1 && mode != 4) return;
    if (tid == 0) s_bad = (mode == 4);
    __syncthreads();
    const uint32_t nb = state[f].n_blocks;
    const uint16_t *sym = out16 + B.out16_off;
    for (uint32_t b = 0; b < nb; b++) {
        __syncthreads();                                           // the previous block's bytes are final; s_bad is settled
        if (s_bad) break;                                          // (uniform: read after the barrier)
        const MfSpecBlock blk = blocks[B.blk_base + b];
        if (blk.kind == MF_BLK_STORED) continue;
        const uint64_t moff = members[(size_t)f * MF_F_MAX_MEMBERS + blk.member].out_off;
        for (uint64_t k = tid; k < blk.out_len; k += 1024) {
            const int v = mf_spec_resolve_sym(sym[blk.out_off + k], B.dst, blk.out_off, moff);
            if (v < 0) s_bad = 1; else B.dst[blk.out_off + k] = (uint8_t)v;
        }
    }
    __syncthreads();
    if (tid == 0 && s_bad) { state[f].mode = 3; kind[B.blob] = 2; status[B.blob] = MF_DEC_E_STREAM; dst_len[B.blob] = 0; }
}

// raw CRC-32 of p[0..n) by a whole CTA (contiguous slices, slice-by-4 from the global tables, GF(2) shifts, xor-reduce)
__device__ uint32_t mf_crc_cta(const uint8_t *p, uint64_t n, uint32_t *s_part) {
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint64_t per = ((n + nt - 1) / nt + 3) & ~3ull;
    const uint64_t lo = min(n, per * tid), hi = min(n, lo + per);
    uint32_t crc = 0;
    uint64_t k = lo;
    for (; k < hi && (((uintptr_t)(p + k)) & 3u); k++) crc = g_mf_crc_tab[0][(crc ^ __ldcg(p + k)) & 0xFFu] ^ (crc >> 8);
    for (; k + 4 <= hi; k += 4) {
        const uint32_t x = crc ^ __ldcg(reinterpret_cast<const uint32_t *>(p + k));
        crc = g_mf_crc_tab[3][x & 0xFFu] ^ g_mf_crc_tab[2][(x >> 8) & 0xFFu] ^ g_mf_crc_tab[1][(x >> 16) & 0xFFu] ^ g_mf_crc_tab[0][x >> 24];
    }
    for (; k < hi; k++) crc = g_mf_crc_tab[0][(crc ^ __ldcg(p + k)) & 0xFFu] ^ (crc >> 8);
    if (hi > lo) crc = mf_gf2_mul(crc, mf_xpow8(n - hi));
    crc = mf_warp_xor(crc);
    if ((tid & 31u) == 0) s_part[tid >> 5] = crc;
    __syncthreads();
    uint32_t r = 0;
    if (tid < 32) { r = tid < (nt >> 5) ? s_part[tid] : 0u; r = mf_warp_xor(r); }
    __syncthreads();
    return r;                                                       // valid in warp 0
}

__global__ void __launch_bounds__(1024)
mf_fverify_kernel(const MfFBlob *__restrict__ fb, const MfSpecMember *__restrict__ members, const MfFState *__restrict__ state,
                  uint8_t *__restrict__ kind, uint64_t *__restrict__ dst_len, int32_t *__restrict__ status) {
    __shared__ uint32_t s_part[32];
    const uint32_t f = blockIdx.x;
    const MfFBlob B = fb[f];
    if (state[f].mode != 1) return;
    const uint32_t nm = state[f].n_members;
    for (uint32_t m = 0; m < nm; m++) {
        const MfSpecMember M = members[(size_t)f * MF_F_MAX_MEMBERS + m];
        const uint32_t raw = mf_crc_cta(B.dst + M.out_off, M.out_len, s_part);
        if (threadIdx.x == 0) {
            int e = 0;
            if (mf_crc_finish(raw, M.out_len) != M.crc_expect) e = MF_DEC_E_CRC;
            else if ((uint32_t)(M.out_len & 0xFFFFFFFFull) != M.isize_expect) e = MF_DEC_E_ISIZE;
            if (e && status[B.blob] == 0) { status[B.blob] = e; dst_len[B.blob] = 0; kind[B.blob] = 2; }
        }
        __syncthreads();
    }
}
