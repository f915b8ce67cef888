// mf_core.h -- format constants and the host/device-shared pieces of the codec:
// CRC-32 GF(2) arithmetic, DEFLATE symbol maps, the CTA-parallel Huffman builder, the
// dynamic-block header writer and the per-lane token emitter.
//
// Everything here is written once and compiled twice: by nvcc for the sm_100a kernels and by
// g++ for the CPU format tests under tests/host_model (MF_PAR_FOR becomes a plain loop).
// The CPU build is test infrastructure; the product only ever runs the CUDA build.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MF_HD __host__ __device__ __forceinline__
#else
#define MF_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define MF_PAR_FOR(t, N) for (int t = (int)threadIdx.x; t < (int)(N); t += (int)blockDim.x)
#define MF_BARRIER() __syncthreads()
#define MF_SERIAL if (threadIdx.x == 0)
#define MF_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define MF_ATOMIC_OR(p, v) atomicOr((p), (v))
#define MF_CLZ(x) __clz((int)(x))
#define MF_BREV(x) __brev(x)
#else
#define MF_PAR_FOR(t, N) for (int t = 0; t < (int)(N); ++t)
#define MF_BARRIER() ((void)0)
#define MF_SERIAL
#define MF_ATOMIC_ADD(p, v) (*(p) += (v))
#define MF_ATOMIC_OR(p, v) (*(p) |= (v))
#define MF_CLZ(x) ((x) ? __builtin_clz((unsigned)(x)) : 32)
static inline uint32_t mf_brev_host(uint32_t v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}
#define MF_BREV(x) mf_brev_host(x)
#endif

// ------------------------------------------------------------------ format constants
// A packed blob is one or more RFC 1952 members.  Each member carries, in its FEXTRA field
// (subfield id 'M','F'), an index that lets the decoder start at every chunk (byte offset) and
// every segment (bit offset) in parallel.  Stock gzip readers skip FEXTRA (RFC 1952 2.3.1.1;
// reference reader /usr/lib/python3.12/gzip.py:463-467).
//
//   chunk   = 32 KiB of raw bytes  = one DEFLATE block (own Huffman tables) followed by an empty
//             stored block (00 00 00 FF FF after bit padding), so every chunk ends byte-aligned.
//   segment = 1 KiB of raw bytes   = the LZ77 window (matches never leave their segment) and the
//             unit one decoder lane decodes; its token bit length is in the index.
//   member  = up to 64 chunks (2 MiB raw: one stripe, the unit one encoder CTA owns); ends with a final empty
//             stored block 01 00 00 FF FF.
#define MF_CHUNK_LOG2 15
#define MF_CHUNK (1u << MF_CHUNK_LOG2)
#define MF_SEG_LOG2 10
#define MF_SEG (1u << MF_SEG_LOG2)
#define MF_SEGS_PER_CHUNK (MF_CHUNK / MF_SEG)       // 32
#define MF_MEMBER_CHUNKS 64u
#define MF_MEMBER_RAW ((uint64_t)MF_MEMBER_CHUNKS * MF_CHUNK)
#define MF_HDR_FIXED 32u                              // 10 gzip + 2 XLEN + 4 subfield hdr + 16 payload hdr
#define MF_IDX_PER_CHUNK 72u                          // 8 (comp_bytes u32, hdr_bits u16, rsv u16) + 32*2 seg bits
#define MF_TRAILER 13u                                // final empty stored block (5) + CRC32 + ISIZE
#define MF_SLOT (MF_CHUNK + 64u)                      // scratch slot per chunk (stored fallback = raw + 5)
#define MF_STORED_MARK 0xFFFFu                        // hdr_bits value marking a stored chunk
#define MF_FORMAT_VERSION 1

#define MF_MIN_MATCH 4
#define MF_MAX_MATCH 258
#define MF_MAXM 128                                   // match-list capacity per segment

MF_HD uint32_t mf_member_header_bytes(uint32_t n_chunks) { return MF_HDR_FIXED + MF_IDX_PER_CHUNK * n_chunks; }
MF_HD uint64_t mf_n_chunks(uint64_t raw) { return (raw + MF_CHUNK - 1) >> MF_CHUNK_LOG2; }
MF_HD uint64_t mf_n_members(uint64_t raw) { return raw == 0 ? 1 : (raw + MF_MEMBER_RAW - 1) / MF_MEMBER_RAW; }
MF_HD uint64_t mf_pack_bound(uint64_t raw) {
    return mf_n_members(raw) * (MF_HDR_FIXED + MF_TRAILER) + mf_n_chunks(raw) * (MF_IDX_PER_CHUNK + 5u) + raw;
}

// ------------------------------------------------------------------ CRC-32 in GF(2)
// Reflected CRC-32 (poly 0xEDB88320).  A 32-bit value v represents the polynomial whose x^0
// coefficient is bit 31.  crc_raw(M) = M(x) * x^32 mod P with zero init and no final xor; the
// standard checksum is crc32(M) = crc_raw(M) ^ (0xFFFFFFFF * x^(8|M|)) ^ 0xFFFFFFFF.
#define MF_CRC_POLY 0xEDB88320u

MF_HD uint32_t mf_gf2_mul(uint32_t a, uint32_t b) {   // a(x)*b(x) mod P (zlib's multmodp, restated)
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & (0x80000000u >> i)) p ^= b;
        b = (b >> 1) ^ ((b & 1u) ? MF_CRC_POLY : 0u);
    }
    return p;
}
// x^(8*nbytes) mod P by square-and-multiply
MF_HD uint32_t mf_xpow8(uint64_t nbytes) {
    uint32_t r = 0x80000000u;        // x^0
    uint32_t sq = 0x00800000u;       // x^8
    while (nbytes) {
        if (nbytes & 1u) r = mf_gf2_mul(r, sq);
        sq = mf_gf2_mul(sq, sq);
        nbytes >>= 1;
    }
    return r;
}
// combine raw CRCs: raw(A||B) = raw(A)*x^(8|B|) ^ raw(B)
MF_HD uint32_t mf_crc_finish(uint32_t raw, uint64_t total_len) {
    return raw ^ mf_gf2_mul(0xFFFFFFFFu, mf_xpow8(total_len)) ^ 0xFFFFFFFFu;
}
// table-free bytewise raw update (used for short tails and by the CPU model)
MF_HD uint32_t mf_crc_raw_byte(uint32_t crc, uint8_t b) {
    crc ^= b;
    for (int k = 0; k < 8; k++) crc = (crc >> 1) ^ ((crc & 1u) ? MF_CRC_POLY : 0u);
    return crc;
}

// ------------------------------------------------------------------ DEFLATE symbol maps (RFC 1951 3.2.5)
MF_HD void mf_len_sym(uint32_t len, uint32_t &sym, uint32_t &ebits, uint32_t &eval) {
    uint32_t l = len - 3;
    if (l < 8) { sym = 257 + l; ebits = 0; eval = 0; return; }
    if (len == 258) { sym = 285; ebits = 0; eval = 0; return; }
    uint32_t k = 31 - MF_CLZ(l);              // >= 3
    sym = 257 + 4 * (k - 1) + ((l >> (k - 2)) & 3u);
    ebits = k - 2;
    eval = l & ((1u << ebits) - 1u);
}
MF_HD void mf_dist_sym(uint32_t dist, uint32_t &sym, uint32_t &ebits, uint32_t &eval) {
    uint32_t d = dist - 1;
    if (d < 4) { sym = d; ebits = 0; eval = 0; return; }
    uint32_t k = 31 - MF_CLZ(d);              // >= 2
    sym = 2 * k + ((d >> (k - 1)) & 1u);
    ebits = k - 1;
    eval = d & ((1u << ebits) - 1u);
}

// match-list entry: pos (10 bits, within segment) | (len-3) << 10 (8 bits) | (dist-1) << 18 (10 bits)
MF_HD uint32_t mf_match_pack(uint32_t pos, uint32_t len, uint32_t dist) { return pos | ((len - 3) << 10) | ((dist - 1) << 18); }
MF_HD uint32_t mf_match_pos(uint32_t m) { return m & 1023u; }
MF_HD uint32_t mf_match_len(uint32_t m) { return ((m >> 10) & 255u) + 3; }
MF_HD uint32_t mf_match_dist(uint32_t m) { return ((m >> 18) & 1023u) + 1; }

// ------------------------------------------------------------------ prefix-code construction (CTA-parallel)
// Code lengths by Kraft-slack refinement instead of a Huffman tree: the tree's two-queue merge is a
// ~300-step serial chain of dependent shared-memory reads (it was 3/4 of the encoder's time on B200),
// whereas this construction is data-parallel except for one short scan per round:
//   1. l_s = ceil(log2(T / f_s)) clamped to [1, maxbits]      (Shannon lengths; Kraft sum <= 1)
//   2. repeat: order the symbols by benefit/cost of shortening them by one bit, f_s * 2^l_s, and
//      shorten greedily while the unused code space (slack) allows.
// Measured against exact Huffman on 32 KiB chunks (float noise, text, low-entropy, skewed): total
// code length within 0.1%.  The result is always a valid (complete or under-full) prefix code.
struct MfHuffScratch {
    uint32_t key[512];       // benefit/cost key per symbol
    uint16_t order[512];     // symbols in greedy order
    uint32_t csort[512];     // cost (code space) of shortening, in greedy order
    uint8_t take[512];       // greedy decision per position
    uint32_t cnt[17];        // codes per length
    uint32_t first[17];      // first canonical code of each length
    uint32_t total;          // sum of counts
    uint32_t mass;           // sum of 2^(maxbits - l) over used symbols
    uint32_t used;           // number of used symbols
    uint32_t again;          // another refinement round is worthwhile
};

#define MF_HUFF_ROUNDS 3

// code word table entry: (len << 16) | bit-reversed code  (ready to OR into an LSB-first stream)
// hist[0..n): symbol counts (modified: fewer than two used symbols are forced to two, as zlib does)
// lens[0..n): code lengths out;  codes[0..n): table entries out.  n <= 512.
template <int NSORT>
MF_HD void mf_huff_build(uint32_t *hist, int n, int maxbits, uint8_t *lens, uint32_t *codes, MfHuffScratch &S) {
    MF_SERIAL { S.total = 0; S.mass = 0; S.used = 0; S.again = 1; for (int l = 0; l <= 16; l++) S.cnt[l] = 0; }
    MF_BARRIER();
    MF_PAR_FOR(s, n) { if (hist[s]) MF_ATOMIC_ADD(&S.used, 1u); }
    MF_BARRIER();
    MF_SERIAL {
        for (int s = 0; S.used < 2 && s < n; s++)
            if (hist[s] == 0) { hist[s] = 1; S.used++; }
    }
    MF_BARRIER();
    MF_PAR_FOR(s, n) { if (hist[s]) MF_ATOMIC_ADD(&S.total, hist[s]); }
    MF_BARRIER();
    const uint32_t full = 1u << maxbits;
    MF_PAR_FOR(s, n) {
        uint32_t f = hist[s], l = 0;
        if (f) {
            uint32_t q = (S.total + f - 1) / f;                  // ceil(T / f)
            l = q <= 1 ? 1u : (uint32_t)(32 - MF_CLZ(q - 1));     // ceil(log2 q)
            if (l < 1) l = 1;
            if (l > (uint32_t)maxbits) l = (uint32_t)maxbits;
            MF_ATOMIC_ADD(&S.mass, full >> l);
        }
        lens[s] = (uint8_t)l;
    }
    MF_BARRIER();
    MF_SERIAL {
        // clamping to maxbits can over-subscribe the code space: lengthen the rarest shortenable symbols
        while (S.mass > full) {
            int best = -1;
            for (int s = 0; s < n; s++)
                if (lens[s] && lens[s] < maxbits && (best < 0 || hist[s] < hist[best])) best = s;
            if (best < 0) break;
            lens[best]++;
            S.mass -= full >> lens[best];
        }
    }
    MF_BARRIER();
    // rounds 0..R-1 shorten every symbol by at most one bit (best code quality); the last round lets a
    // symbol take several bits, which always uses up the remaining code space: DEFLATE decoders
    // (zlib's inflate_table) reject an under-full code, so the Kraft sum must end at exactly 1.
    for (int round = 0; round <= MF_HUFF_ROUNDS; round++) {
        const bool last = (round == MF_HUFF_ROUNDS) || !S.again;
        MF_BARRIER();
        MF_PAR_FOR(t, n) { uint32_t l = lens[t]; S.key[t] = (l > 1) ? (hist[t] << l) : 0u; }   // hist <= 2^15, l <= 15
        MF_BARRIER();
        MF_PAR_FOR(t, n) {
            uint32_t k = S.key[t], r = 0;
            for (int u = 0; u < n; u++) { uint32_t ku = S.key[u]; r += (ku > k) || (ku == k && u < t); }
            S.order[r] = (uint16_t)t;
            S.csort[r] = k ? ((full >> lens[t]) | ((uint32_t)lens[t] << 24)) : 0u;     // cost | length << 24
        }
        MF_BARRIER();
        MF_SERIAL {
            uint32_t slack = full - S.mass, changed = 0;
            int i = 0;
            for (; i < n && slack; i++) {
                uint32_t e = S.csort[i];
                if (!e) break;
                uint32_t c = e & 0xFFFFFFu, l = e >> 24, tk = 0;
                while (l > 1 && c <= slack && (last || tk == 0)) { slack -= c; c <<= 1; l--; tk++; }
                S.take[i] = (uint8_t)tk;
                changed |= tk;
            }
            for (; i < n; i++) S.take[i] = 0;
            S.mass = full - slack;
            S.again = changed && slack;
        }
        MF_BARRIER();
        MF_PAR_FOR(i, n) { if (S.take[i]) lens[S.order[i]] -= S.take[i]; }
        MF_BARRIER();
        if (last) break;
    }
    MF_BARRIER();
    MF_PAR_FOR(s, n) { if (lens[s]) MF_ATOMIC_ADD(&S.cnt[lens[s]], 1u); }
    MF_BARRIER();
    MF_SERIAL {
        uint32_t code = 0;
        S.cnt[0] = 0;
        for (int l = 1; l <= maxbits; l++) { code = (code + S.cnt[l - 1]) << 1; S.first[l] = code; }
    }
    MF_BARRIER();
    MF_PAR_FOR(s, n) {                                   // canonical code = first[len] + rank among equal lengths
        uint32_t l = lens[s];
        uint32_t e = 0;
        if (l) {
            uint32_t r = 0;
            for (int u = 0; u < s; u++) r += (lens[u] == l);
            uint32_t code = S.first[l] + r;
            e = (l << 16) | (MF_BREV(code) >> (32 - l));
        }
        codes[s] = e;
    }
    MF_BARRIER();
}

// ------------------------------------------------------------------ LSB-first bit writer over 32-bit words
// The target words must be zero-initialised.  The first and the last word a writer touches may
// be shared with its neighbours and are OR-ed atomically; interior words are plain stores.
struct MfBitWriter {
    uint32_t *out;
    uint32_t wi;       // current word index
    uint64_t acc;      // pending bits, LSB first
    uint32_t nacc;     // number of pending bits (the low `start shift` bits of the first word are not ours)
    bool first;
    MF_HD void init(uint32_t *o, uint32_t bitpos) { out = o; wi = bitpos >> 5; acc = 0; nacc = bitpos & 31u; first = true; }
    MF_HD void put(uint32_t v, uint32_t nbits) {         // nbits <= 32
        acc |= (uint64_t)v << nacc;
        nacc += nbits;
        if (nacc >= 32) {
            uint32_t w = (uint32_t)acc;
            if (first) { MF_ATOMIC_OR(&out[wi], w); first = false; } else out[wi] = w;
            wi++; acc >>= 32; nacc -= 32;
        }
    }
    MF_HD void flush() { if (nacc) MF_ATOMIC_OR(&out[wi], (uint32_t)acc); }
};

// ------------------------------------------------------------------ dynamic block header (RFC 1951 3.2.7)
struct MfHdrScratch {
    uint8_t lens[320];       // litlen lengths followed by dist lengths
    uint8_t rle_sym[320];
    uint8_t rle_ext[320];
    uint32_t clhist[19];
    uint8_t cllens[19];
    uint32_t clcodes[19];
    uint32_t nrle, hlit, hdist, hclen;
    uint32_t hdr_bits;
};

// Serial part 1: trim trailing zero lengths, run-length encode with symbols 16/17/18, histogram.
MF_HD void mf_hdr_rle(const uint8_t *ll_lens, const uint8_t *d_lens, MfHdrScratch &H) {
    int hlit = 286; while (hlit > 257 && ll_lens[hlit - 1] == 0) hlit--;
    int hdist = 30; while (hdist > 1 && d_lens[hdist - 1] == 0) hdist--;
    H.hlit = (uint32_t)hlit; H.hdist = (uint32_t)hdist;
    int n = hlit + hdist;
    for (int i = 0; i < hlit; i++) H.lens[i] = ll_lens[i];
    for (int i = 0; i < hdist; i++) H.lens[hlit + i] = d_lens[i];
    for (int i = 0; i < 19; i++) H.clhist[i] = 0;
    int i = 0, r = 0;
    while (i < n) {
        uint8_t v = H.lens[i];
        int run = 1;
        while (i + run < n && H.lens[i + run] == v) run++;
        if (v == 0 && run >= 3) {
            int take = run > 138 ? 138 : run;
            if (take >= 11) { H.rle_sym[r] = 18; H.rle_ext[r] = (uint8_t)(take - 11); }
            else { H.rle_sym[r] = 17; H.rle_ext[r] = (uint8_t)(take - 3); }
            H.clhist[H.rle_sym[r]]++; r++; i += take;
        } else if (v != 0 && run >= 4) {
            H.rle_sym[r] = v; H.rle_ext[r] = 0; H.clhist[v]++; r++; i++;   // the value itself
            int rest = run - 1;
            while (rest >= 3) {
                int take = rest > 6 ? 6 : rest;
                H.rle_sym[r] = 16; H.rle_ext[r] = (uint8_t)(take - 3); H.clhist[16]++; r++;
                i += take; rest -= take;
            }
        } else {
            H.rle_sym[r] = v; H.rle_ext[r] = 0; H.clhist[v]++; r++; i++;
        }
    }
    H.nrle = (uint32_t)r;
}

// Serial part 2 (after the code-length Huffman code is built): write BFINAL/BTYPE + tables.
// Returns the number of header bits written starting at bit 0 of `out`.
MF_HD uint32_t mf_hdr_write(uint32_t *out, MfHdrScratch &H) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19; while (hclen > 4 && H.cllens[order[hclen - 1]] == 0) hclen--;
    H.hclen = (uint32_t)hclen;
    MfBitWriter w; w.init(out, 0);
    w.put(0u, 1);                 // BFINAL = 0 (the member ends with a separate final empty block)
    w.put(2u, 2);                 // BTYPE = 10 dynamic
    w.put(H.hlit - 257, 5); w.put(H.hdist - 1, 5); w.put((uint32_t)hclen - 4, 4);
    for (int i = 0; i < hclen; i++) w.put(H.cllens[order[i]], 3);
    for (uint32_t i = 0; i < H.nrle; i++) {
        uint32_t s = H.rle_sym[i];
        uint32_t e = H.clcodes[s];
        w.put(e & 0xFFFFu, e >> 16);
        if (s == 16) w.put(H.rle_ext[i], 2);
        else if (s == 17) w.put(H.rle_ext[i], 3);
        else if (s == 18) w.put(H.rle_ext[i], 7);
    }
    uint32_t bits = (w.wi << 5) + w.nacc;
    w.flush();
    H.hdr_bits = bits;
    return bits;
}

// ------------------------------------------------------------------ per-lane token emission
// A segment's tokens are the literals of its raw bytes with the (sorted, non-overlapping) match
// list substituted in.  Lane `lane` of 32 owns the tokens that START in bytes [32*lane, 32*lane+32)
// of the segment.  With w == nullptr the function only counts bits (sweep 1); otherwise it also
// writes them (sweep 2).  `seg` points at the segment's raw bytes, `seglen` <= 1024.
MF_HD uint32_t mf_lane_emit(const uint8_t *seg, uint32_t seglen, const uint32_t *matches, uint32_t nmatch,
                            uint32_t lane, const uint32_t *ll_codes, const uint32_t *d_codes, MfBitWriter *w) {
    uint32_t lo = lane * 32u, hi = lo + 32u;
    if (hi > seglen) hi = seglen;
    if (lo >= hi) return 0;
    // first match whose start is >= lo (binary search), and the end of the one before it
    uint32_t a = 0, b = nmatch;
    while (a < b) { uint32_t mid = (a + b) >> 1; if (mf_match_pos(matches[mid]) < lo) a = mid + 1; else b = mid; }
    uint32_t mi = a;
    uint32_t p = lo;
    if (mi > 0) { uint32_t pm = matches[mi - 1]; uint32_t e = mf_match_pos(pm) + mf_match_len(pm); if (e > p) p = e; }
    uint32_t bits = 0;
    uint32_t next_pos = mi < nmatch ? mf_match_pos(matches[mi]) : 0xFFFFFFFFu;
    while (p < hi) {
        if (p == next_pos) {
            uint32_t m = matches[mi];
            uint32_t len = mf_match_len(m), dist = mf_match_dist(m);
            uint32_t ls, lb, lv, ds, db, dv;
            mf_len_sym(len, ls, lb, lv); mf_dist_sym(dist, ds, db, dv);
            uint32_t le = ll_codes[ls], de = d_codes[ds];
            bits += (le >> 16) + lb + (de >> 16) + db;
            if (w) {
                w->put((le & 0xFFFFu) | (lv << (le >> 16)), (le >> 16) + lb);     // <= 15 + 5 bits
                w->put((de & 0xFFFFu) | (dv << (de >> 16)), (de >> 16) + db);     // <= 15 + 9 bits
            }
            p += len; mi++;
            next_pos = mi < nmatch ? mf_match_pos(matches[mi]) : 0xFFFFFFFFu;
        } else {
            uint32_t e = ll_codes[seg[p]];
            bits += e >> 16;
            if (w) w->put(e & 0xFFFFu, e >> 16);
            p++;
        }
    }
    return bits;
}
