// mf_inflate_core.h -- serial (one-lane) RFC 1951 / RFC 1952 decoding, shared by the CUDA build
// (one GPU lane per foreign stream; block-header parsing for the chunk-parallel decoder) and by
// the g++ format tests.  Replaces zlib inflate behind ContentAddressedStore._unpack_v1
// (metaflow/datastore/content_addressed_store.py:218-220; gzip.py:446-480, 520-566).
// Memory-safe on arbitrary input: every read and write is bounds-checked.
#pragma once
#include "mf_core.h"

#define MF_DEC_OK 0
#define MF_DEC_E_DST 1       // == MFCAS_E_DST_TOO_SMALL
#define MF_DEC_E_STREAM 3    // == MFCAS_E_BAD_STREAM
#define MF_DEC_E_CRC 4       // == MFCAS_E_BAD_CRC
#define MF_DEC_E_ISIZE 5     // == MFCAS_E_BAD_ISIZE
#define MF_DEC_E_TRUNC 8     // == MFCAS_E_TRUNC

// LSB-first bit reader over a byte array (byte loads: works for any alignment and address space)
struct MfBitReader {
    const uint8_t *p;
    uint64_t n;        // bytes available
    uint64_t pos;      // next byte to load
    uint64_t buf;
    uint32_t cnt;      // valid bits in buf
    bool over;         // tried to read past the end
    MF_HD void init(const uint8_t *src, uint64_t nbytes, uint64_t bitpos) {
        p = src; n = nbytes; pos = bitpos >> 3; buf = 0; cnt = 0; over = false;
        uint32_t skip = (uint32_t)(bitpos & 7u);
        if (skip) { refill(); buf >>= skip; cnt = cnt >= skip ? cnt - skip : 0; }
    }
    MF_HD void refill() {
        // an aligned 32-bit word at a time where possible (a serial decoder pays every byte load's latency)
        if (cnt <= 32 && pos + 4 <= n && (((uintptr_t)(p + pos)) & 3u) == 0) {
            buf |= (uint64_t)(*reinterpret_cast<const uint32_t *>(p + pos)) << cnt; cnt += 32; pos += 4;
        }
        while (cnt <= 56 && pos < n) { buf |= (uint64_t)p[pos++] << cnt; cnt += 8; }
    }
    MF_HD uint32_t peek(uint32_t k) { if (cnt < k) refill(); return (uint32_t)(buf & ((1ull << k) - 1ull)); }
    MF_HD void drop(uint32_t k) { if (k > cnt) { over = true; buf = 0; cnt = 0; } else { buf >>= k; cnt -= k; } }
    MF_HD uint32_t get(uint32_t k) { if (!k) return 0; uint32_t v = peek(k); drop(k); return v; }
    MF_HD uint64_t bitpos() const { return pos * 8ull - cnt; }
};

// canonical Huffman decoding tables (count per length + symbols sorted by (length, symbol))
struct MfCanon {
    uint16_t count[16];
    uint16_t symbol[288];
};

// returns 0 complete, >0 incomplete (unused code space), <0 over-subscribed
MF_HD int mf_canon_build(MfCanon &h, const uint8_t *len, int n) {
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

MF_HD int mf_canon_decode(MfBitReader &br, const MfCanon &h) {
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

// 9-bit first-level table beside the canonical one, for streams decoded by a single lane (reference-written
// files have no index, so a whole blob is one serial stream): one lookup per symbol instead of a bit-by-bit walk
// over up to 15 lengths, each step a dependent local-memory load.  entry = symbol | length << 9; 0 = longer code.
#define MF_GEN_LUT_BITS 9
struct MfCanonLut { uint16_t e[1u << MF_GEN_LUT_BITS]; };
MF_HD void mf_canon_lut_build(MfCanonLut &t, const MfCanon &h, const uint8_t *len, int n) {
    for (uint32_t i = 0; i < (1u << MF_GEN_LUT_BITS); i++) t.e[i] = 0;
    uint32_t next[16];
    uint32_t code = 0;
    next[0] = 0;
    for (int l = 1; l < 16; l++) { code = (code + (l > 1 ? h.count[l - 1] : 0)) << 1; next[l] = code; }
    for (int s = 0; s < n; s++) {
        const uint32_t l = len[s] & 15u;
        if (!l) continue;
        const uint32_t c = next[l]++;
        if (l > MF_GEN_LUT_BITS) continue;
        uint32_t rev = 0;
        for (uint32_t k = 0; k < l; k++) rev |= ((c >> k) & 1u) << (l - 1 - k);
        const uint16_t ent = (uint16_t)((uint32_t)s | (l << 9));
        for (uint32_t k = rev; k < (1u << MF_GEN_LUT_BITS); k += (1u << l)) t.e[k] = ent;
    }
}
MF_HD int mf_canon_decode_lut(MfBitReader &br, const MfCanon &h, const MfCanonLut &t) {
    const uint32_t bits = br.peek(15);
    const uint32_t ent = t.e[bits & ((1u << MF_GEN_LUT_BITS) - 1u)];
    if (ent) { br.drop(ent >> 9); return br.over ? -1 : (int)(ent & 511u); }
    return mf_canon_decode(br, h);
}

struct MfDeflateTables {
    uint16_t lbase[29]; uint8_t lext[29]; uint16_t dbase[30]; uint8_t dext[30];
};
MF_HD uint32_t mf_lbase(int s) { return s < 8 ? 3u + s : (s == 28 ? 258u : ((4u + (s & 3)) << ((s >> 2) - 1)) + 3u); }
MF_HD uint32_t mf_lext(int s) { return (s < 8 || s == 28) ? 0u : (uint32_t)((s >> 2) - 1); }
MF_HD uint32_t mf_dbase(int s) { return s < 4 ? 1u + s : ((2u + (s & 1)) << ((s >> 1) - 1)) + 1u; }
MF_HD uint32_t mf_dext(int s) { return s < 4 ? 0u : (uint32_t)((s >> 1) - 1); }

// Parse a dynamic block's tables (after BFINAL/BTYPE were consumed): lens[0..hlit) litlen lengths,
// lens[hlit..hlit+hdist) distance lengths.  Returns 0 or MF_DEC_E_*.
MF_HD int mf_read_dynamic_tables(MfBitReader &br, uint8_t *lens, uint32_t &hlit, uint32_t &hdist) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    hlit = br.get(5) + 257; hdist = br.get(5) + 1;
    uint32_t hclen = br.get(4) + 4;
    if (br.over) return MF_DEC_E_TRUNC;
    if (hlit > 286 || hdist > 30) return MF_DEC_E_STREAM;
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (uint32_t i = 0; i < hclen; i++) cl[order[i]] = (uint8_t)br.get(3);
    if (br.over) return MF_DEC_E_TRUNC;
    MfCanon ch;
    if (mf_canon_build(ch, cl, 19) != 0) return MF_DEC_E_STREAM;
    uint32_t i = 0, tot = hlit + hdist;
    while (i < tot) {
        int sym = mf_canon_decode(br, ch);
        if (sym < 0) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        uint8_t v = 0; uint32_t rep;
        if (sym == 16) { if (i == 0) return MF_DEC_E_STREAM; v = lens[i - 1]; rep = 3 + br.get(2); }
        else if (sym == 17) rep = 3 + br.get(3);
        else rep = 11 + br.get(7);
        if (br.over) return MF_DEC_E_TRUNC;
        if (i + rep > tot) return MF_DEC_E_STREAM;
        while (rep--) lens[i++] = v;
    }
    if (lens[256] == 0) return MF_DEC_E_STREAM;
    return 0;
}

// The same parse (BFINAL/BTYPE included), the quick way, for a reader R with a 64-bit buffer and word refills
// (fill(): more than 32 valid bits afterwards; peek/drop/get without checks; used(): bits consumed; reads past the end
// of the blob deliver zeros): a 7-bit first-level table for the code-length code instead of the bit-by-bit canonical walk
// over a byte reader, which cost the chunk decoder almost half of what decoding the chunk's 32 segments does.
// Returns the header's length in bits, or 0 for ANYTHING unusual -- the caller then runs the parser above, which owns the
// error codes.  When it returns non-zero, lens / hlit / hdist are exactly what the parser above delivers
// (tests/test_format_model.py::test_fast_header_parser_equals_the_portable_one).  cl_lut: 128 bytes of scratch.
MF_HD uint32_t mf_bitrev32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
template <typename R>
MF_HD uint32_t mf_read_header_fast_t(R &br, uint32_t comp_bytes, uint8_t *lens, uint8_t *cl_lut, uint32_t &hlit_out, uint32_t &hdist_out) {
    br.fill();
    if (br.get(3) != 4u) return 0;                        // BFINAL 0, BTYPE 10
    const uint32_t hlit = br.get(5) + 257u, hdist = br.get(5) + 1u, hclen = br.get(4) + 4u;
    if (hlit > 286u || hdist > 30u) return 0;
    // code-length code: lengths in the RFC's order; a complete code is required (mf_canon_build(...) == 0 with a code)
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t cl[19];
    uint32_t cnt[8], next[8];
    for (int k = 0; k < 19; k++) cl[k] = 0;
    for (int k = 0; k < 8; k++) cnt[k] = 0;
    for (uint32_t i = 0; i < hclen; i++) {
        if ((i & 7u) == 0) br.fill();                     // 8 x 3 bits per refill
        const uint32_t l = br.get(3);
        cl[order[i]] = (uint8_t)l;
        cnt[l]++;
    }
    int left = 1;
    uint32_t code = 0;
    next[0] = 0;
    for (int l = 1; l < 8; l++) { code = (code + (l > 1 ? cnt[l - 1] : 0u)) << 1; next[l] = code; left = (left << 1) - (int)cnt[l]; }
    if (left != 0) return 0;
    for (uint32_t s = 0; s < 19u; s++) {
        const uint32_t l = cl[s];
        if (!l) continue;
        const uint32_t c = next[l]++;
        const uint32_t rev = mf_bitrev32(c) >> (32u - l);
        const uint8_t ent = (uint8_t)(s | (l << 5));
        for (uint32_t k = rev; k < 128u; k += (1u << l)) cl_lut[k] = ent;
    }
    const uint32_t tot = hlit + hdist;
    uint32_t i = 0, prev = 0;
    while (i < tot) {
        br.fill();                                        // > 32 bits: a code (<= 7) and its extra bits (<= 7)
        const uint32_t e = cl_lut[br.peek(7)];
        const uint32_t sym = e & 31u;
        br.drop(e >> 5);
        if (sym < 16u) { lens[i++] = (uint8_t)sym; prev = sym; continue; }
        uint32_t v = 0, rep;
        if (sym == 16u) { if (i == 0) return 0; v = prev; rep = 3u + br.get(2); }
        else if (sym == 17u) rep = 3u + br.get(3);
        else rep = 11u + br.get(7);
        if (i + rep > tot) return 0;
        for (uint32_t k = 0; k < rep; k++) lens[i + k] = (uint8_t)v;
        i += rep; prev = v;
    }
    if (lens[256] == 0) return 0;
    const uint32_t used = br.used();
    if ((uint64_t)used > (uint64_t)comp_bytes * 8ull) return 0;
    hlit_out = hlit; hdist_out = hdist;
    return used;
}

// Address arithmetic of the chunk decoder's input ring (MfRingReader, mf_inflate.cuh), kept here so that the CPU model runs
// the very expressions the kernel runs.  A warp's ring is MF_RING_BYTES = 4 slots x 32 lanes x 16 bytes, MF_RING_BYTES-aligned
// in the shared address space; lane l's quad in slot s sits at s * 512 + l * 16.
#define MF_RING_BYTES 2048u
// read pointer after the word at `raddr`: +4 inside a quad; from a quad's last word to the same lane's quad in the next slot
// (+500 inside the ring's 11 address bits, summed BEFORE it can carry out of them: lane 31's quad ends at the slot's last byte)
MF_HD uint32_t mf_ring_next_word(uint32_t raddr) {
    return (raddr & 15u) == 12u ? ((raddr & ~(MF_RING_BYTES - 1u)) | ((raddr + 500u) & (MF_RING_BYTES - 1u))) : raddr + 4u;
}
// write pointer after a quad went to `wslot`: the same lane's quad in the next slot
MF_HD uint32_t mf_ring_next_slot(uint32_t wslot) { return (wslot & ~(MF_RING_BYTES - 1u)) | ((wslot + 512u) & (MF_RING_BYTES - 1u)); }
// slots behind the one being read that have not been refilled yet
MF_HD uint32_t mf_ring_free_slots(uint32_t raddr, uint32_t wslot) { return ((raddr - wslot) >> 9) & 3u; }

// Whole raw-DEFLATE stream, one lane.  dst doubles as the 32 KiB window.
// tables of one serial inflater.  The caller decides where they live: on the GPU a per-thread (local-memory) copy is
// interleaved with the other threads' at a 128-byte stride -- a 1 KiB table is 256 cache lines and every lookup
// misses L1 -- so the foreign-stream kernel keeps them in shared memory.
struct MfInflateWork { MfCanon ll, dd; MfCanonLut lut; uint8_t lens[320]; };

MF_HD int mf_inflate_stream(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *consumed, uint64_t *produced, MfInflateWork &work) {
    MfBitReader br; br.init(src, n, 0);
    uint64_t op = 0;
    uint32_t last = 0;
    MfCanon &ll = work.ll, &dd = work.dd;
    MfCanonLut &lut = work.lut;
    uint8_t *lens = work.lens;
    while (!last) {
        last = br.get(1);
        uint32_t type = br.get(2);
        if (br.over) return MF_DEC_E_TRUNC;
        if (type == 0) {
            uint64_t bp = (br.bitpos() + 7ull) >> 3;          // skip to the byte boundary
            if (bp + 4 > n) return MF_DEC_E_TRUNC;
            uint32_t len = src[bp] | ((uint32_t)src[bp + 1] << 8);
            uint32_t nlen = src[bp + 2] | ((uint32_t)src[bp + 3] << 8);
            if ((len ^ 0xFFFFu) != nlen) return MF_DEC_E_STREAM;
            bp += 4;
            if (bp + len > n) return MF_DEC_E_TRUNC;
            if (op + len > cap) return MF_DEC_E_DST;
            for (uint32_t k = 0; k < len; k++) dst[op + k] = src[bp + k];
            op += len;
            br.init(src, n, (bp + len) * 8ull);
            continue;
        }
        if (type == 3) return MF_DEC_E_STREAM;
        if (type == 1) {
            int s = 0;
            for (; s < 144; s++) lens[s] = 8;
            for (; s < 256; s++) lens[s] = 9;
            for (; s < 280; s++) lens[s] = 7;
            for (; s < 288; s++) lens[s] = 8;
            mf_canon_build(ll, lens, 288);
            mf_canon_lut_build(lut, ll, lens, 288);
            for (s = 0; s < 30; s++) lens[s] = 5;
            mf_canon_build(dd, lens, 30);
        } else {
            uint32_t hlit, hdist;
            int e = mf_read_dynamic_tables(br, lens, hlit, hdist);
            if (e) return e;
            int r = mf_canon_build(ll, lens, (int)hlit);
            if (r < 0 || (r > 0 && (int)hlit - ll.count[0] != 1)) return MF_DEC_E_STREAM;
            mf_canon_lut_build(lut, ll, lens, (int)hlit);
            r = mf_canon_build(dd, lens + hlit, (int)hdist);
            if (r < 0 || (r > 0 && (int)hdist - dd.count[0] != 1)) return MF_DEC_E_STREAM;
        }
        for (;;) {
            int sym = mf_canon_decode_lut(br, ll, lut);
            if (sym < 0) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
            if (sym < 256) {
                if (op >= cap) return MF_DEC_E_DST;
                dst[op++] = (uint8_t)sym;
            } else if (sym == 256) {
                break;
            } else {
                sym -= 257;
                if (sym >= 29) return MF_DEC_E_STREAM;
                uint32_t len = mf_lbase(sym) + br.get(mf_lext(sym));
                int ds = mf_canon_decode(br, dd);
                if (ds < 0) return br.over ? MF_DEC_E_TRUNC : MF_DEC_E_STREAM;
                if (ds >= 30) return MF_DEC_E_STREAM;
                uint32_t dist = mf_dbase(ds) + br.get(mf_dext(ds));
                if (br.over) return MF_DEC_E_TRUNC;
                if (dist > op) return MF_DEC_E_STREAM;
                if (op + len > cap) return MF_DEC_E_DST;
                for (uint32_t k = 0; k < len; k++) { dst[op] = dst[op - dist]; op++; }
            }
        }
    }
    if (br.over) return MF_DEC_E_TRUNC;
    *consumed = (br.bitpos() + 7ull) >> 3;
    *produced = op;
    return 0;
}

// RFC 1952 member header length (>= 10) or 0 on error; *err set to MF_DEC_E_*.
MF_HD uint64_t mf_gzip_header_len(const uint8_t *p, uint64_t n, int *err) {
    *err = 0;
    if (n < 10) { *err = MF_DEC_E_TRUNC; return 0; }
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8) { *err = MF_DEC_E_STREAM; return 0; }
    uint8_t flg = p[3];
    uint64_t pos = 10;
    if (flg & 4) { if (pos + 2 > n) { *err = MF_DEC_E_TRUNC; return 0; } uint64_t xl = p[pos] | ((uint32_t)p[pos + 1] << 8); pos += 2 + xl; }
    if (flg & 8) { while (pos < n && p[pos]) pos++; pos++; }
    if (flg & 16) { while (pos < n && p[pos]) pos++; pos++; }
    if (flg & 2) pos += 2;
    if (pos > n) { *err = MF_DEC_E_TRUNC; return 0; }
    return pos;
}

MF_HD uint32_t mf_le32(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
MF_HD uint32_t mf_le16(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8); }

// Is the member at p an 'MF' indexed member?  On success fills raw_len / n_chunks and returns the
// header length (index included).
MF_HD uint64_t mf_is_indexed_member(const uint8_t *p, uint64_t n, uint32_t *raw_len, uint32_t *n_chunks) {
    if (n < MF_HDR_FIXED) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || p[3] != 4) return 0;
    if (p[12] != 'M' || p[13] != 'F') return 0;
    if (p[16] != MF_FORMAT_VERSION || p[17] != MF_CHUNK_LOG2 || p[18] != MF_SEG_LOG2) return 0;
    uint32_t rl = mf_le32(p + 20), nc = mf_le32(p + 24);
    if (nc > 768u || nc != (uint32_t)mf_n_chunks(rl)) return 0;      // writers: 64 chunks per member (768 in format revision 0)
    uint64_t hdr = mf_member_header_bytes(nc);
    if (mf_le16(p + 10) != hdr - 12 || mf_le16(p + 14) != hdr - 16 || hdr > n) return 0;
    *raw_len = rl; *n_chunks = nc;
    return hdr;
}

// Full gunzip of one blob by one lane (any RFC 1952 writer): members until the input ends,
// CRC-32 + ISIZE of each verified, zero padding between / after members skipped.
// crc_tab: 256-entry reflected CRC-32 table.
MF_HD int mf_gunzip_stream(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *produced, const uint32_t *crc_tab) {
    uint64_t ip = 0, op = 0;
    int members = 0;
    while (ip < n) {
        if (members) { while (ip < n && src[ip] == 0) ip++; if (ip >= n) break; }
        int err;
        uint64_t hl = mf_gzip_header_len(src + ip, n - ip, &err);
        if (!hl) return err;
        ip += hl;
        uint64_t used = 0, made = 0;
        MfInflateWork work;
        int e = mf_inflate_stream(src + ip, n - ip, dst + op, cap - op, &used, &made, work);
        if (e) return e;
        ip += used;
        if (ip + 8 > n) return MF_DEC_E_TRUNC;
        uint32_t crc = 0xFFFFFFFFu;
        for (uint64_t k = 0; k < made; k++) crc = crc_tab[(crc ^ dst[op + k]) & 0xFFu] ^ (crc >> 8);
        crc = ~crc;
        if (crc != mf_le32(src + ip)) return MF_DEC_E_CRC;
        if ((uint32_t)(made & 0xFFFFFFFFull) != mf_le32(src + ip + 4)) return MF_DEC_E_ISIZE;
        ip += 8; op += made; members++;
    }
    if (!members) return MF_DEC_E_TRUNC;
    *produced = op;
    return 0;
}
