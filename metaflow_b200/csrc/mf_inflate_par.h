// mf_inflate_par.h -- block-parallel decoding of FOREIGN gzip streams (stock Metaflow / any RFC 1952 writer: one
// DEFLATE stream per member, no index), host/device shared like mf_inflate_core.h (the g++ build is the CPU model the
// tests run; the product only runs the CUDA build).
//
// Replaces the serial walk of ContentAddressedStore._unpack_v1 (metaflow/datastore/content_addressed_store.py:218-220,
// gzip.py:446-566 -> zlib inflate) for files THIS library did not write.  A DEFLATE stream is bit-serial only between
// blocks; inside the stream every dynamic-Huffman block starts with a self-describing header, so:
//
//   1. scan    every bit position is tested for "a valid dynamic block header starts here" (BTYPE, HLIT/HDIST ranges,
//              a COMPLETE code-length code, a well-formed length sequence with an end-of-block code, complete
//              literal/length and distance codes -- the checks zlib's inflate_table makes); false positives are rare
//              and harmless (step 3 never reaches them)
//   2. size    every candidate is decoded by its own lane WITHOUT output: where its block ends, how many bytes it makes
//   3. walk    one thread follows the true chain of a blob -- member header, then block after block: a stored block is
//              skipped by its LEN, a dynamic block must be a sized candidate (else the blob falls back to the serial
//              decoder), a fixed block is sized on the spot -- giving every block its output offset
//   4. decode  every block of the chain is decoded again, now into 16-bit symbols at its output offset: a byte, or a
//              MARKER 0x8000 | (w - 1) for "the byte w positions before this block's first byte" (a match that reaches
//              into earlier blocks; copies of markers stay markers)
//   5. resolve blocks in order: markers are replaced by the (already final) bytes they name
//   6. verify  CRC-32 and ISIZE of every member
#pragma once
#include "mf_inflate_core.h"

#define MF_SPEC_MARK 0x8000u
#define MF_SPEC_F_VALID 1u
#define MF_SPEC_F_FINAL 2u

// block kinds of the walked chain
#define MF_BLK_STORED 0u
#define MF_BLK_FIXED 1u
#define MF_BLK_DYNAMIC 2u

struct MfSpecCand {            // a position that parses as a dynamic block header, sized by step 2
    uint64_t start_bit;        // of the BFINAL bit, relative to the blob's first byte
    uint64_t end_bit;          // first bit after the end-of-block code
    uint64_t out_len;          // bytes the block produces
    uint32_t flags;            // MF_SPEC_F_*
    uint32_t pad;
};

struct MfSpecBlock {           // a block of the true chain (step 3)
    uint64_t start_bit;        // BFINAL bit (stored: the first raw byte * 8)
    uint64_t out_off;          // offset of its first byte in the blob's decoded output
    uint64_t out_len;
    uint32_t kind;             // MF_BLK_*
    uint32_t member;           // index of the member it belongs to
};

struct MfSpecMember {          // a gzip member of a foreign blob (step 3), verified in step 6
    uint64_t out_off, out_len;
    uint32_t crc_expect, isize_expect;
};

// tables of one block decoder (2 KiB: one per lane in shared memory)
struct MfCanonDist { uint16_t count[16]; uint16_t symbol[32]; };
struct MfSpecWork { MfCanon ll; MfCanonDist dd; MfCanonLut lut; uint8_t lens[320]; };

MF_HD int mf_canon_build_dist(MfCanonDist &h, const uint8_t *len, int n) {
    uint16_t offs[16];
    for (int i = 0; i < 16; i++) h.count[i] = 0;
    for (int s = 0; s < n; s++) h.count[len[s] & 15]++;
    if (h.count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l < 16; l++) { left <<= 1; left -= h.count[l]; if (left < 0) return left; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + h.count[l]);
    for (int s = 0; s < n; s++) if (len[s]) h.symbol[offs[len[s]]++] = (uint16_t)s;
    return left;
}
MF_HD int mf_canon_decode_dist(MfBitReader &br, const MfCanonDist &h) {
    uint32_t bits = br.peek(15);
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int)(bits & 1u); bits >>= 1;
        int cnt = h.count[l];
        if (code - cnt < first) { br.drop((uint32_t)l); return br.over ? -1 : h.symbol[index + (code - first)]; }
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return -1;
}

// Fixed-Huffman tables (RFC 1951 3.2.6)
MF_HD void mf_spec_fixed_tables(MfSpecWork &W) {
    int s = 0;
    for (; s < 144; s++) W.lens[s] = 8;
    for (; s < 256; s++) W.lens[s] = 9;
    for (; s < 280; s++) W.lens[s] = 7;
    for (; s < 288; s++) W.lens[s] = 8;
    mf_canon_build(W.ll, W.lens, 288);
    mf_canon_lut_build(W.lut, W.ll, W.lens, 288);
    for (s = 0; s < 30; s++) W.lens[s] = 5;
    mf_canon_build_dist(W.dd, W.lens, 30);
}

// Tables of the dynamic block whose BFINAL bit is at `start_bit`; br is left at the first symbol.
// Applies the validity rules of zlib's inflate (complete codes, or a single code of length 1; an end-of-block code).
MF_HD int mf_spec_dynamic_tables(MfBitReader &br, MfSpecWork &W, uint32_t *final_flag) {
    const uint32_t last = br.get(1), type = br.get(2);
    if (br.over) return MF_DEC_E_TRUNC;
    if (type != 2) return MF_DEC_E_STREAM;
    uint32_t hlit, hdist;
    int e = mf_read_dynamic_tables(br, W.lens, hlit, hdist);
    if (e) return e;
    int r = mf_canon_build(W.ll, W.lens, (int)hlit);
    if (r < 0 || (r > 0 && (int)hlit - W.ll.count[0] != 1)) return MF_DEC_E_STREAM;
    mf_canon_lut_build(W.lut, W.ll, W.lens, (int)hlit);
    r = mf_canon_build_dist(W.dd, W.lens + hlit, (int)hdist);
    if (r < 0 || (r > 0 && (int)hdist - W.dd.count[0] != 1)) return MF_DEC_E_STREAM;
    *final_flag = last;
    return 0;
}

// Step 1, the cheap part: can a dynamic block header start at this bit?  `w` holds the 17 bits from the position on
// (LSB first): BFINAL, BTYPE (2), HLIT (5), HDIST (5), HCLEN (4).
MF_HD bool mf_spec_quick(uint32_t w) {
    if (((w >> 1) & 3u) != 2u) return false;
    if (((w >> 3) & 31u) > 29u) return false;           // HLIT <= 286 - 257
    if (((w >> 8) & 31u) > 29u) return false;           // HDIST <= 30 - 1
    return true;
}

// Step 1, the code-length code: HCLEN+4 lengths of 3 bits each must form a COMPLETE prefix code (zlib rejects anything
// else for this code).  `bits` = the 57 bits after the 17 fixed ones.
MF_HD bool mf_spec_clcode_complete(uint32_t hclen4, uint64_t bits) {
    uint32_t sum = 0;                                    // Kraft sum in units of 2^-7
    for (uint32_t i = 0; i < hclen4; i++) {
        const uint32_t l = (uint32_t)(bits >> (3u * i)) & 7u;
        if (l) sum += 128u >> l;
    }
    return sum == 128u;
}

// Step 1, the full check of one position that passed the cheap tests: the code-length sequence must be well formed
// and the two codes it describes must be usable (zlib's rules: complete, or a single code of length 1; an end-of-block
// code).  Needs ~400 bytes of scratch (on the GPU: local memory; it runs for ~0.2 % of the bit positions).
MF_HD bool mf_spec_header_valid(const uint8_t *src, uint64_t n, uint64_t start_bit) {
    MfBitReader br; br.init(src, n, start_bit + 3);
    uint8_t lens[320];
    uint32_t hlit, hdist;
    if (mf_read_dynamic_tables(br, lens, hlit, hdist) != 0) return false;
    for (int pass = 0; pass < 2; pass++) {
        const uint8_t *l = pass ? lens + hlit : lens;
        const int cnt = pass ? (int)hdist : (int)hlit;
        int count[16];
        for (int i = 0; i < 16; i++) count[i] = 0;
        for (int s = 0; s < cnt; s++) count[l[s] & 15]++;
        if (count[0] == cnt) { if (pass == 0) return false; continue; }      // no distance codes at all is fine
        int left = 1;
        for (int k = 1; k < 16; k++) { left <<= 1; left -= count[k]; if (left < 0) return false; }
        if (left > 0 && cnt - count[0] != 1) return false;
    }
    return true;
}

// Step 1 + 2 for one candidate position: parse the header, decode to the end-of-block code, produce nothing.
// max_out: a block that would make more output than this is not one (bounds the work spent on a false positive).
MF_HD int mf_spec_size_block(const uint8_t *src, uint64_t n, uint64_t start_bit, uint64_t max_out, MfSpecWork &W, MfSpecCand *out) {
    MfBitReader br; br.init(src, n, start_bit);
    uint32_t last = 0;
    int e = mf_spec_dynamic_tables(br, W, &last);
    if (e) return e;
    uint64_t op = 0;
    for (;;) {
        int sym = mf_canon_decode_lut(br, W.ll, W.lut);
        if (sym < 0) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        if (sym < 256) { op++; if (op > max_out) return MF_DEC_E_DST; continue; }
        if (sym == 256) break;
        sym -= 257;
        if (sym >= 29) return MF_DEC_E_STREAM;
        const uint32_t len = mf_lbase(sym) + br.get(mf_lext(sym));
        const int ds = mf_canon_decode_dist(br, W.dd);
        if (ds < 0) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        if (ds >= 30) return MF_DEC_E_STREAM;
        const uint32_t dist = mf_dbase(ds) + br.get(mf_dext(ds));
        if (br.over) return MF_DEC_E_TRUNC;
        if (dist > op + 32768u) return MF_DEC_E_STREAM;       // reaches further back than any window allows
        op += len;
        if (op > max_out) return MF_DEC_E_DST;
    }
    if (br.over) return MF_DEC_E_TRUNC;
    out->start_bit = start_bit; out->end_bit = br.bitpos(); out->out_len = op;
    out->flags = MF_SPEC_F_VALID | (last ? MF_SPEC_F_FINAL : 0u);
    out->pad = 0;
    return 0;
}

// Size of a fixed-Huffman block (walker, on the spot): br is behind BFINAL/BTYPE.
MF_HD int mf_spec_size_fixed(MfBitReader &br, uint64_t max_out, MfSpecWork &W, uint64_t *out_len) {
    mf_spec_fixed_tables(W);
    uint64_t op = 0;
    for (;;) {
        int sym = mf_canon_decode_lut(br, W.ll, W.lut);
        if (sym < 0) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        if (sym < 256) { op++; if (op > max_out) return MF_DEC_E_DST; continue; }
        if (sym == 256) break;
        sym -= 257;
        if (sym >= 29) return MF_DEC_E_STREAM;
        const uint32_t len = mf_lbase(sym) + br.get(mf_lext(sym));
        const int ds = mf_canon_decode_dist(br, W.dd);
        if (ds < 0 || ds >= 30) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        const uint32_t dist = mf_dbase(ds) + br.get(mf_dext(ds));
        if (br.over) return MF_DEC_E_TRUNC;
        if (dist > op + 32768u) return MF_DEC_E_STREAM;
        op += len;
        if (op > max_out) return MF_DEC_E_DST;
    }
    *out_len = op;
    return 0;
}

// Step 4 for one block of the chain: decode into 16-bit symbols out16[0 .. blk.out_len).  Stored blocks are copied as
// plain bytes.  Matches that reach before the block's first byte leave markers.
MF_HD int mf_spec_decode_block(const uint8_t *src, uint64_t n, const MfSpecBlock &blk, MfSpecWork &W, uint16_t *out16) {
    if (blk.kind == MF_BLK_STORED) {
        const uint64_t bp = blk.start_bit >> 3;
        if (bp + blk.out_len > n) return MF_DEC_E_TRUNC;
        for (uint64_t k = 0; k < blk.out_len; k++) out16[k] = src[bp + k];
        return 0;
    }
    MfBitReader br; br.init(src, n, blk.start_bit);
    if (blk.kind == MF_BLK_FIXED) {
        br.get(3);
        mf_spec_fixed_tables(W);
    } else {
        uint32_t last;
        int e = mf_spec_dynamic_tables(br, W, &last);
        if (e) return e;
    }
    uint64_t op = 0;
    const uint64_t cap = blk.out_len;
    for (;;) {
        int sym = mf_canon_decode_lut(br, W.ll, W.lut);
        if (sym < 0) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        if (sym < 256) {
            if (op >= cap) return MF_DEC_E_STREAM;
            out16[op++] = (uint16_t)sym;
            continue;
        }
        if (sym == 256) break;
        sym -= 257;
        if (sym >= 29) return MF_DEC_E_STREAM;
        const uint32_t len = mf_lbase(sym) + br.get(mf_lext(sym));
        const int ds = mf_canon_decode_dist(br, W.dd);
        if (ds < 0 || ds >= 30) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        const uint32_t dist = mf_dbase(ds) + br.get(mf_dext(ds));
        if (br.over) return MF_DEC_E_TRUNC;
        if (op + len > cap) return MF_DEC_E_STREAM;
        for (uint32_t k = 0; k < len; k++) {
            uint16_t v;
            if (dist <= op) v = out16[op - dist];
            else {
                const uint64_t w = dist - op;                 // that many bytes before the block's first byte
                if (w > 32768u) return MF_DEC_E_STREAM;
                v = (uint16_t)(MF_SPEC_MARK | (uint32_t)(w - 1u));
            }
            out16[op++] = v;
        }
    }
    return op == cap ? 0 : MF_DEC_E_STREAM;
}

// Step 3: follow the chain of one blob.  cands[0..n_cands) sorted by start_bit.  Writes blocks[] / members[] (up to the
// given capacities) and the total decoded size.  Returns 0, or an MF_DEC_E_* for a stream that is definitely bad, or
// -1 when the chain cannot be followed with what steps 1-2 found (caller falls back to the serial decoder).
#define MF_SPEC_FALLBACK (-1)
MF_HD int mf_spec_walk(const uint8_t *src, uint64_t n, const MfSpecCand *cands, uint32_t n_cands, uint64_t out_cap, MfSpecWork &W,
                       MfSpecBlock *blocks, uint32_t max_blocks, uint32_t *n_blocks, MfSpecMember *members, uint32_t max_members,
                       uint32_t *n_members, uint64_t *total_out) {
    uint64_t ip = 0, op = 0;
    uint32_t nb = 0, nm = 0, ci = 0;
    while (ip < n) {
        if (nm) { while (ip < n && src[ip] == 0) ip++; if (ip >= n) break; }
        int err;
        const uint64_t hl = mf_gzip_header_len(src + ip, n - ip, &err);
        if (!hl) return err;
        ip += hl;
        if (nm >= max_members) return MF_SPEC_FALLBACK;
        const uint64_t member_out0 = op;
        uint64_t bit = ip * 8ull;
        uint32_t last = 0;
        while (!last) {
            MfBitReader br; br.init(src, n, bit);
            last = br.get(1);
            const uint32_t type = br.get(2);
            if (br.over) return MF_DEC_E_TRUNC;
            if (type == 3) return MF_DEC_E_STREAM;
            if (nb >= max_blocks) return MF_SPEC_FALLBACK;
            MfSpecBlock &B = blocks[nb];
            B.member = nm; B.out_off = op;
            if (type == 0) {
                const uint64_t bp = (bit + 3ull + 7ull) >> 3;
                if (bp + 4 > n) return MF_DEC_E_TRUNC;
                const uint32_t len = src[bp] | ((uint32_t)src[bp + 1] << 8), nlen = src[bp + 2] | ((uint32_t)src[bp + 3] << 8);
                if ((len ^ 0xFFFFu) != nlen) return MF_DEC_E_STREAM;
                if (bp + 4 + len > n) return MF_DEC_E_TRUNC;
                B.kind = MF_BLK_STORED; B.start_bit = (bp + 4) * 8ull; B.out_len = len;
                bit = (bp + 4 + len) * 8ull;
            } else if (type == 1) {
                uint64_t made = 0;
                const int e = mf_spec_size_fixed(br, out_cap - op, W, &made);
                if (e) return e;
                B.kind = MF_BLK_FIXED; B.start_bit = bit; B.out_len = made;
                bit = br.bitpos();
            } else {
                while (ci < n_cands && cands[ci].start_bit < bit) ci++;
                if (ci >= n_cands || cands[ci].start_bit != bit || !(cands[ci].flags & MF_SPEC_F_VALID)) return MF_SPEC_FALLBACK;
                B.kind = MF_BLK_DYNAMIC; B.start_bit = bit; B.out_len = cands[ci].out_len;
                bit = cands[ci].end_bit;
            }
            if (B.out_len > out_cap - op) { *total_out = op + B.out_len; return MF_DEC_E_DST; }
            // a match may not reach before its member's first byte: checked in resolve (markers) and here for the sizes
            op += B.out_len;
            nb++;
        }
        ip = (bit + 7ull) >> 3;
        if (ip + 8 > n) return MF_DEC_E_TRUNC;
        MfSpecMember &M = members[nm];
        M.out_off = member_out0; M.out_len = op - member_out0;
        M.crc_expect = mf_le32(src + ip); M.isize_expect = mf_le32(src + ip + 4);
        ip += 8; nm++;
    }
    if (!nm) return MF_DEC_E_TRUNC;
    *n_blocks = nb; *n_members = nm; *total_out = op;
    return 0;
}

// Step 5 for one symbol of a block: the final byte.  dst = the blob's decoded output (bytes before blk.out_off are final),
// member_off = out_off of the block's member (a marker may not reach before it).  Returns -1 for an invalid reference.
MF_HD int mf_spec_resolve_sym(uint16_t s, const uint8_t *dst, uint64_t blk_out_off, uint64_t member_off) {
    if (!(s & MF_SPEC_MARK)) return (int)(s & 0xFFu);
    const uint64_t w = (uint64_t)(s & 0x7FFFu) + 1u;
    if (w > blk_out_off - member_off) return -1;
    return dst[blk_out_off - w];
}
