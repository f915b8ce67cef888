// CPU model of the MF container format (TEST INFRASTRUCTURE ONLY -- never part of the product).
//
// Compiles the host/device-shared headers of the codec (metaflow_b200/csrc/mf_core.h,
// mf_inflate_core.h) with g++ and drives them the way the CUDA kernels do, one "lane" at a
// time, so that the Huffman builder, block-header writer, bit packer, GF(2) CRC arithmetic,
// member layout and the serial inflater can be checked against zlib on a machine without a GPU.
// The LZ77 stage here is a plain sequential greedy matcher (the warp-parallel one exists only
// in CUDA); it produces the same kind of match list (sorted, confined to 1 KiB segments).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../metaflow_b200/csrc/mf_core.h"
#include "../../metaflow_b200/csrc/mf_inflate_core.h"
#include "../../metaflow_b200/csrc/mf_inflate_par.h"

static uint32_t crc_tab[256];
static void init_tab() {
    static bool done = false;
    if (done) return;
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? MF_CRC_POLY : 0u); crc_tab[i] = c; }
    done = true;
}

struct ChunkOut { std::vector<uint8_t> bytes; uint32_t hdr_bits; uint16_t seg_bits[MF_SEGS_PER_CHUNK]; uint32_t crc_raw; };

static void encode_chunk(const uint8_t *src, uint32_t len, bool lz, ChunkOut &out) {
    init_tab();
    std::vector<uint8_t> data(MF_CHUNK + 32, 0);
    memcpy(data.data(), src, len);
    // CRC: 64-byte pieces combined with GF(2) multipliers, exactly as the kernel does
    uint32_t crc = 0;
    for (uint32_t g = 0; g < 512; g++) {
        uint32_t start = g * 64;
        if (start >= len) break;
        uint32_t end = start + 64 < len ? start + 64 : len;
        uint32_t c = 0;
        for (uint32_t p = start; p < end; p++) c = crc_tab[(c ^ data[p]) & 0xFF] ^ (c >> 8);
        crc ^= mf_gf2_mul(c, mf_xpow8(len - end));
    }
    out.crc_raw = crc;
    static uint32_t ll_hist[288], d_hist[32], ll_codes[288], d_codes[32];
    static uint8_t ll_lens[288], d_lens[32];
    memset(ll_hist, 0, sizeof ll_hist); memset(d_hist, 0, sizeof d_hist);
    std::vector<std::vector<uint32_t>> matches(MF_SEGS_PER_CHUNK);
    for (uint32_t seg = 0; seg < MF_SEGS_PER_CHUNK; seg++) {
        uint32_t base = seg * MF_SEG;
        if (base >= len) break;
        uint32_t seglen = len - base < MF_SEG ? len - base : MF_SEG;
        std::vector<int> head(1 << 10, -1);
        uint32_t p = 0;
        while (p < seglen) {
            bool took = false;
            if (lz && p + MF_MIN_MATCH <= seglen && matches[seg].size() < MF_MAXM) {
                uint32_t w; memcpy(&w, &data[base + p], 4);
                uint32_t h = (w * 2654435761u) >> 22;
                int q = head[h];
                head[h] = (int)p;
                if (q >= 0) {
                    uint32_t maxlen = seglen - p < MF_MAX_MATCH ? seglen - p : MF_MAX_MATCH, l = 0;
                    while (l < maxlen && data[base + q + l] == data[base + p + l]) l++;
                    uint32_t dist = p - (uint32_t)q;
                    if (l >= 6 || (l >= 4 && dist <= 32) || (l == 5 && dist <= 256)) {
                        matches[seg].push_back(mf_match_pack(p, l, dist));
                        uint32_t ls, lb, lv, ds, db, dv;
                        mf_len_sym(l, ls, lb, lv); mf_dist_sym(dist, ds, db, dv);
                        ll_hist[ls]++; d_hist[ds]++;
                        p += l; took = true;
                    }
                }
            }
            if (!took) { ll_hist[data[base + p]]++; p++; }
        }
    }
    ll_hist[256] = 1;
    static MfHuffScratch HS; static MfHdrScratch HD;
    mf_huff_build<512>(ll_hist, 286, 15, ll_lens, ll_codes, HS);
    mf_huff_build<32>(d_hist, 30, 15, d_lens, d_codes, HS);
    mf_hdr_rle(ll_lens, d_lens, HD);
    mf_huff_build<32>(HD.clhist, 19, 7, HD.cllens, HD.clcodes, HS);
    std::vector<uint32_t> words(MF_CHUNK / 4 + 1024, 0);
    mf_hdr_write(words.data(), HD);
    uint32_t lane_bits[MF_SEGS_PER_CHUNK][32];
    uint32_t pos = HD.hdr_bits;
    uint32_t seg_off[MF_SEGS_PER_CHUNK];
    for (uint32_t seg = 0; seg < MF_SEGS_PER_CHUNK; seg++) {
        uint32_t base = seg * MF_SEG, tot = 0;
        for (uint32_t lane = 0; lane < 32; lane++) {
            uint32_t b = 0;
            if (base < len) {
                uint32_t seglen = len - base < MF_SEG ? len - base : MF_SEG;
                b = mf_lane_emit(data.data() + base, seglen, matches[seg].data(), (uint32_t)matches[seg].size(), lane, ll_codes, d_codes, nullptr);
            }
            lane_bits[seg][lane] = b; tot += b;
        }
        out.seg_bits[seg] = (uint16_t)tot; seg_off[seg] = pos; pos += tot;
    }
    uint32_t eob = ll_codes[256];
    uint32_t end_bits = pos + (eob >> 16) + 3;
    uint32_t dyn_bytes = ((end_bits + 7) >> 3) + 4;
    if (dyn_bytes >= len + 5) {
        out.bytes.resize(len + 5);
        out.bytes[0] = 0; out.bytes[1] = (uint8_t)len; out.bytes[2] = (uint8_t)(len >> 8); out.bytes[3] = (uint8_t)~len; out.bytes[4] = (uint8_t)(~len >> 8);
        memcpy(out.bytes.data() + 5, src, len);
        out.hdr_bits = MF_STORED_MARK;
        memset(out.seg_bits, 0, sizeof out.seg_bits);
        return;
    }
    // lanes in a scrambled order: the result must not depend on who writes first
    for (uint32_t k = 0; k < MF_SEGS_PER_CHUNK * 32; k++) {
        uint32_t j = (k * 7919u) % (MF_SEGS_PER_CHUNK * 32);
        uint32_t seg = j / 32, lane = j % 32, base = seg * MF_SEG;
        if (base >= len || !lane_bits[seg][lane]) continue;
        uint32_t off = seg_off[seg];
        for (uint32_t l = 0; l < lane; l++) off += lane_bits[seg][l];
        uint32_t seglen = len - base < MF_SEG ? len - base : MF_SEG;
        MfBitWriter bw; bw.init(words.data(), off);
        mf_lane_emit(data.data() + base, seglen, matches[seg].data(), (uint32_t)matches[seg].size(), lane, ll_codes, d_codes, &bw);
        bw.flush();
    }
    MfBitWriter bw; bw.init(words.data(), pos); bw.put(eob & 0xFFFF, eob >> 16); bw.flush();
    uint32_t byte = ((end_bits + 7) >> 3) + 2;
    uint8_t *ob = reinterpret_cast<uint8_t *>(words.data());
    ob[byte] = 0xFF; ob[byte + 1] = 0xFF;
    out.bytes.assign(ob, ob + dyn_bytes);
    out.hdr_bits = HD.hdr_bits;
}

static void put16(std::vector<uint8_t> &v, size_t o, uint32_t x) { v[o] = (uint8_t)x; v[o + 1] = (uint8_t)(x >> 8); }
static void put32(std::vector<uint8_t> &v, size_t o, uint32_t x) { for (int i = 0; i < 4; i++) v[o + i] = (uint8_t)(x >> (8 * i)); }

extern "C" {

uint64_t model_pack_bound(uint64_t n) { return mf_pack_bound(n); }

// returns packed length (0 on overflow of cap)
uint64_t model_pack(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, int lz) {
    std::vector<uint8_t> out;
    uint64_t nm = mf_n_members(n);
    for (uint64_t m = 0; m < nm; m++) {
        uint64_t off = m * MF_MEMBER_RAW;
        uint32_t rl = (uint32_t)(n - off < MF_MEMBER_RAW ? n - off : MF_MEMBER_RAW);
        uint32_t nc = (uint32_t)mf_n_chunks(rl);
        std::vector<ChunkOut> ch(nc);
        uint32_t payload = 0;
        for (uint32_t c = 0; c < nc; c++) {
            uint32_t cl = rl - c * MF_CHUNK < MF_CHUNK ? rl - c * MF_CHUNK : MF_CHUNK;
            encode_chunk(src + off + (uint64_t)c * MF_CHUNK, cl, lz != 0, ch[c]);
            payload += (uint32_t)ch[c].bytes.size();
        }
        uint32_t hdr = mf_member_header_bytes(nc);
        size_t base = out.size();
        out.resize(base + hdr + payload + MF_TRAILER, 0);
        uint8_t *h = out.data() + base;
        h[0] = 0x1f; h[1] = 0x8b; h[2] = 8; h[3] = 4; h[9] = 255;
        put16(out, base + 10, hdr - 12); h[12] = 'M'; h[13] = 'F'; put16(out, base + 14, hdr - 16);
        h[16] = MF_FORMAT_VERSION; h[17] = MF_CHUNK_LOG2; h[18] = MF_SEG_LOG2; h[19] = 0;
        put32(out, base + 20, rl); put32(out, base + 24, nc);
        size_t p = base + hdr;
        uint32_t crc = 0;
        for (uint32_t c = 0; c < nc; c++) {
            put32(out, base + MF_HDR_FIXED + 8 * c, (uint32_t)ch[c].bytes.size());
            put16(out, base + MF_HDR_FIXED + 8 * c + 4, ch[c].hdr_bits);
            for (uint32_t s = 0; s < MF_SEGS_PER_CHUNK; s++) put16(out, base + MF_HDR_FIXED + 8 * nc + 64 * c + 2 * s, ch[c].seg_bits[s]);
            memcpy(out.data() + p, ch[c].bytes.data(), ch[c].bytes.size());
            p += ch[c].bytes.size();
            uint64_t end = (uint64_t)(c + 1) * MF_CHUNK < rl ? (uint64_t)(c + 1) * MF_CHUNK : rl;
            crc ^= mf_gf2_mul(ch[c].crc_raw, mf_xpow8(rl - end));
        }
        out[p] = 1; out[p + 1] = 0; out[p + 2] = 0; out[p + 3] = 0xFF; out[p + 4] = 0xFF;
        put32(out, p + 5, mf_crc_finish(crc, rl)); put32(out, p + 9, rl);
    }
    if (out.size() > cap) return 0;
    memcpy(dst, out.data(), out.size());
    return out.size();
}

int model_gunzip(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *produced) {
    init_tab();
    return mf_gunzip_stream(src, n, dst, cap, produced, crc_tab);
}

// Decode an MF member the way the chunk kernel does: every segment independently from its indexed
// bit offset (serial reader, canonical decoding).  Returns 0 or MF_DEC_E_*.
int model_unpack_indexed(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *produced) {
    init_tab();
    uint64_t ip = 0, op = 0;
    int members = 0;
    while (ip < n) {
        if (members) { while (ip < n && src[ip] == 0) ip++; if (ip >= n) break; }
        uint32_t rl, nc;
        uint64_t hdr = mf_is_indexed_member(src + ip, n - ip, &rl, &nc);
        if (!hdr) return MF_DEC_E_STREAM;
        if (op + rl > cap) return MF_DEC_E_DST;
        const uint8_t *h = src + ip;
        uint64_t coff = hdr;
        uint32_t crc = 0;
        for (uint32_t c = 0; c < nc; c++) {
            uint32_t cb = mf_le32(h + MF_HDR_FIXED + 8 * c), hb = mf_le16(h + MF_HDR_FIXED + 8 * c + 4);
            uint32_t cl = rl - c * MF_CHUNK < MF_CHUNK ? rl - c * MF_CHUNK : MF_CHUNK;
            const uint8_t *cs = h + coff;
            uint8_t *out = dst + op + (uint64_t)c * MF_CHUNK;
            if (hb == MF_STORED_MARK) {
                if (cb != cl + 5) return MF_DEC_E_STREAM;
                memcpy(out, cs + 5, cl);
            } else {
                MfBitReader br; br.init(cs, cb, 0);
                if (br.get(1) != 0 || br.get(2) != 2) return MF_DEC_E_STREAM;
                uint8_t lens[320]; uint32_t hlit, hdist;
                int e = mf_read_dynamic_tables(br, lens, hlit, hdist);
                if (e) return e;
                if (br.bitpos() != hb) return MF_DEC_E_STREAM;
                MfCanon ll, dd;
                mf_canon_build(ll, lens, (int)hlit); mf_canon_build(dd, lens + hlit, (int)hdist);
                uint64_t bit = hb;
                for (uint32_t s = 0; s < MF_SEGS_PER_CHUNK; s++) {
                    uint32_t sb = mf_le16(h + MF_HDR_FIXED + 8 * nc + 64 * c + 2 * s);
                    uint32_t base = s * MF_SEG;
                    if (base >= cl) { if (sb) return MF_DEC_E_STREAM; continue; }
                    uint32_t seglen = cl - base < MF_SEG ? cl - base : MF_SEG;
                    MfBitReader sr; sr.init(cs, cb, bit);
                    uint32_t o = 0;
                    while (o < seglen) {
                        int sym = mf_canon_decode(sr, ll);
                        if (sym < 0 || sym == 256) return MF_DEC_E_STREAM;
                        if (sym < 256) { out[base + o++] = (uint8_t)sym; continue; }
                        sym -= 257;
                        uint32_t len = mf_lbase(sym) + sr.get(mf_lext(sym));
                        int ds = mf_canon_decode(sr, dd);
                        if (ds < 0) return MF_DEC_E_STREAM;
                        uint32_t dist = mf_dbase(ds) + sr.get(mf_dext(ds));
                        if (dist > o || o + len > seglen) return MF_DEC_E_STREAM;
                        for (uint32_t k = 0; k < len; k++) { out[base + o] = out[base + o - dist]; o++; }
                    }
                    if (sr.bitpos() != bit + sb) return MF_DEC_E_STREAM;
                    bit += sb;
                }
            }
            uint32_t cc = 0;
            for (uint32_t k = 0; k < cl; k++) cc = crc_tab[(cc ^ out[k]) & 0xFF] ^ (cc >> 8);
            uint64_t end = (uint64_t)c * MF_CHUNK + cl;
            crc ^= mf_gf2_mul(cc, mf_xpow8(rl - end));
            coff += cb;
        }
        const uint8_t *t = h + coff;
        if (mf_crc_finish(crc, rl) != mf_le32(t + 5)) return MF_DEC_E_CRC;
        if (mf_le32(t + 9) != rl) return MF_DEC_E_ISIZE;
        ip += coff + MF_TRAILER; op += rl; members++;
    }
    *produced = op;
    return members ? 0 : MF_DEC_E_TRUNC;
}

// Block-parallel decoding of a foreign gzip blob, the way the CUDA kernels do it (mf_inflate_par.h), one "lane" at a
// time: scan every bit position, size the candidates, walk the chain, decode into 16-bit symbols, resolve, verify.
// Returns 0 / MF_DEC_E_*, or -1 when the chain could not be followed (the product then uses the serial decoder).
// stats[0] = candidates found, stats[1] = blocks of the chain, stats[2] = markers resolved.
static uint64_t peek_bits(const uint8_t *src, uint64_t n, uint64_t bitpos, uint32_t count) {
    uint64_t v = 0;
    for (uint32_t k = 0; k < count; k++) {
        const uint64_t b = bitpos + k;
        if ((b >> 3) >= n) break;
        v |= (uint64_t)((src[b >> 3] >> (b & 7u)) & 1u) << k;
    }
    return v;
}

int model_gunzip_parallel(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *produced, uint64_t *stats) {
    init_tab();
    static MfSpecWork W;
    std::vector<MfSpecCand> cands;
    for (uint64_t p = 0; p + 17 <= n * 8ull; p++) {
        const uint32_t w = (uint32_t)peek_bits(src, n, p, 17);
        if (!mf_spec_quick(w)) continue;
        if (!mf_spec_clcode_complete(((w >> 13) & 15u) + 4u, peek_bits(src, n, p + 17, 57))) continue;
        if (!mf_spec_header_valid(src, n, p)) continue;
        MfSpecCand c;
        if (mf_spec_size_block(src, n, p, cap, W, &c) == 0) cands.push_back(c);
    }
    std::vector<MfSpecBlock> blocks(cands.size() + n / 5 + 16);
    std::vector<MfSpecMember> members(n / 18 + 2);
    uint32_t nb = 0, nm = 0;
    uint64_t total = 0;
    int e = mf_spec_walk(src, n, cands.data(), (uint32_t)cands.size(), cap, W, blocks.data(), (uint32_t)blocks.size(), &nb,
                         members.data(), (uint32_t)members.size(), &nm, &total);
    if (stats) { stats[0] = cands.size(); stats[1] = nb; stats[2] = 0; }
    if (e) { if (e == MF_DEC_E_DST) *produced = total; return e; }
    std::vector<uint16_t> out16(total + 1);
    for (uint32_t b = 0; b < nb; b++) {
        e = mf_spec_decode_block(src, n, blocks[b], W, out16.data() + blocks[b].out_off);
        if (e) return e;
    }
    for (uint32_t b = 0; b < nb; b++) {
        const MfSpecBlock &B = blocks[b];
        for (uint64_t k = 0; k < B.out_len; k++) {
            const uint16_t s = out16[B.out_off + k];
            if (stats && (s & MF_SPEC_MARK)) stats[2]++;
            const int v = mf_spec_resolve_sym(s, dst, B.out_off, members[B.member].out_off);
            if (v < 0) return MF_DEC_E_STREAM;
            dst[B.out_off + k] = (uint8_t)v;
        }
    }
    for (uint32_t m = 0; m < nm; m++) {
        uint32_t crc = 0xFFFFFFFFu;
        for (uint64_t k = 0; k < members[m].out_len; k++) crc = crc_tab[(crc ^ dst[members[m].out_off + k]) & 0xFFu] ^ (crc >> 8);
        if (~crc != members[m].crc_expect) return MF_DEC_E_CRC;
        if ((uint32_t)(members[m].out_len & 0xFFFFFFFFull) != members[m].isize_expect) return MF_DEC_E_ISIZE;
    }
    *produced = total;
    return 0;
}

// ---- the chunk decoder's quick block-header parser against the portable one (both from mf_inflate_core.h)
struct HostWordReader {
    const uint8_t *p; uint64_t n, pos; uint64_t buf; uint32_t cnt, fetched;
    uint32_t word() { uint32_t v = 0; for (int k = 0; k < 4; k++) if (pos + k < n) v |= (uint32_t)p[pos + k] << (8 * k); pos += 4; return v; }
    void init(const uint8_t *src, uint64_t nbytes) { p = src; n = nbytes; pos = 0; buf = word(); buf |= (uint64_t)word() << 32; cnt = 64; fetched = 64; }
    void fill() { if (cnt <= 32u) { buf |= (uint64_t)word() << cnt; cnt += 32u; fetched += 32u; } }
    uint32_t peek(uint32_t k) { return (uint32_t)buf & ((1u << k) - 1u); }
    void drop(uint32_t k) { buf >>= k; cnt -= k; }
    uint32_t get(uint32_t k) { uint32_t v = peek(k); drop(k); return v; }
    uint32_t used() const { return fetched - cnt; }
};

// out[0..2] = fast: bits (0 = declined), hlit, hdist; out[3..6] = portable: error, bits, hlit, hdist.
// `n` bytes are readable (the blob), the chunk is the first comp_bytes of them.
void model_header_both(const uint8_t *src, uint64_t n, uint32_t comp_bytes, uint8_t *lens_fast, uint8_t *lens_port, uint32_t *out) {
    uint8_t cl_lut[128];
    memset(cl_lut, 0, sizeof cl_lut);
    uint32_t hl = 0, hd = 0;
    HostWordReader wr; wr.init(src, n);
    out[0] = mf_read_header_fast_t(wr, comp_bytes, lens_fast, cl_lut, hl, hd);
    out[1] = hl; out[2] = hd;
    MfBitReader br; br.init(src, comp_bytes, 0);
    uint32_t bfinal = br.get(1), btype = br.get(2);
    int e = 0;
    hl = hd = 0;
    if (bfinal != 0 || btype != 2) e = MF_DEC_E_STREAM;
    if (!e) e = mf_read_dynamic_tables(br, lens_port, hl, hd);
    out[3] = (uint32_t)e; out[4] = e ? 0u : (uint32_t)br.bitpos(); out[5] = hl; out[6] = hd;
}

// ---- the chunk decoder's input-ring address arithmetic (the expressions MfRingReader runs on the GPU)
uint32_t model_ring_next_word(uint32_t raddr) { return mf_ring_next_word(raddr); }
uint32_t model_ring_next_slot(uint32_t wslot) { return mf_ring_next_slot(wslot); }
uint32_t model_ring_free_slots(uint32_t raddr, uint32_t wslot) { return mf_ring_free_slots(raddr, wslot); }

}  // extern "C"
