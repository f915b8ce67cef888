"""CPU checks of the codec's host/device-shared code (mf_core.h, mf_inflate_core.h) compiled with
g++: Huffman builder, block-header writer, bit packer, GF(2) CRC arithmetic, member layout and the
serial inflater, validated against zlib / gzip and the C oracle."""
import gzip
import os
import zlib

import numpy as np
import pytest

import model_lib
from oracle import c_oracle


def _cases():
    rng = np.random.default_rng(0)
    return {
        "empty": b"",
        "one": b"a",
        "abc": b"abc" * 3,
        "zeros": bytes(200000),
        "normal": rng.standard_normal(70000, dtype=np.float32).tobytes(),
        "uniform": rng.random(30000, dtype=np.float32).tobytes(),
        "lowent": rng.integers(0, 16, 60000).astype(np.float32).tobytes(),
        "text": b"The quick brown fox jumps over the lazy dog. " * 2000,
        "random": rng.bytes(100001),
        "ramp": bytes(range(256)) * 300,
        "seg_edge": rng.bytes(1024) + b"x" * 1024 + rng.bytes(7),
        "skew": bytes([0] * 60000 + list(range(256))),
        "chunk_minus1": rng.bytes(32767),
        "chunk_exact": rng.integers(0, 4, 32768).astype(np.uint8).tobytes(),
        "chunk_plus1": rng.integers(0, 4, 32769).astype(np.uint8).tobytes(),
    }


@pytest.mark.parametrize("lz", [0, 1])
def test_model_stream_is_valid_gzip(lz):
    for name, blob in _cases().items():
        z = model_lib.pack(blob, lz)
        assert len(z) <= model_lib.lib().model_pack_bound(len(blob)), name
        assert gzip.decompress(z) == blob, name            # CPython gzip = what the reference's reader runs
        e, out = c_oracle.gunzip(z, len(blob))             # independent RFC 1951/1952 restatement
        assert e == 0 and out == blob, (name, e)
        e, out = model_lib.gunzip(z, len(blob))            # the product's serial inflater (foreign path)
        assert e == 0 and out == blob, (name, e)
        e, out = model_lib.unpack_indexed(z, len(blob))    # segment-parallel decoding order
        assert e == 0 and out == blob, (name, e)


def test_model_ratio_not_worse_than_zlib_level1_on_float_noise():
    rng = np.random.default_rng(3)
    for blob in (rng.standard_normal(200000, dtype=np.float32).tobytes(), rng.random(200000, dtype=np.float32).tobytes()):
        ours = len(model_lib.pack(blob, 1))
        assert ours <= len(gzip.compress(blob, 1)) * 1.005


def test_serial_inflater_on_zlib_streams():
    for name, blob in _cases().items():
        for level in (0, 1, 3, 9):
            z = gzip.compress(blob, level)
            e, out = model_lib.gunzip(z + z + b"\0" * 5, 2 * len(blob))
            assert e == 0 and out == blob + blob, (name, level, e)
    co = zlib.compressobj(9, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)   # fixed-Huffman blocks
    data = b"abcabcabc" * 500
    z = co.compress(data) + co.flush()
    e, out = model_lib.gunzip(z, len(data))
    assert e == 0 and out == data


def test_serial_inflater_long_codes_and_every_symbol():
    """Streams whose literal/length codes are longer than the inflater's 9-bit first-level table (skewed
    distributions give 10..15-bit codes to rare symbols) and that use long matches and far distances."""
    rng = np.random.default_rng(77)
    skew = rng.choice(256, size=300000, p=np.array([0.5 ** min(i + 1, 22) for i in range(256)]) / sum(0.5 ** min(i + 1, 22) for i in range(256))).astype(np.uint8).tobytes()
    every = bytes(range(256)) * 40 + bytes(200000) + bytes(rng.integers(0, 256, 70000, dtype=np.uint8)) + (b"0123456789abcdef" * 4096)
    text = (open(gzip.__file__, "rb").read() * 6)
    for name, blob in (("skew", skew), ("every", every), ("text", text)):
        for level in (1, 6, 9):
            z = gzip.compress(blob, level)
            e, out = model_lib.gunzip(z, len(blob))
            assert e == 0 and out == blob, (name, level, e)
            # truncated and bit-flipped streams are reported, never decoded to the wrong bytes silently
            e, out = model_lib.gunzip(z[: len(z) // 2], len(blob))
            assert e != 0
            bad = bytearray(z); bad[len(z) // 2] ^= 0x10
            e, out = model_lib.gunzip(bytes(bad), len(blob))
            assert e != 0 or out != blob


def test_multi_member_boundary():
    rng = np.random.default_rng(5)
    n = 768 * 32768                       # MF_MEMBER_RAW
    for size in (n - 1, n, n + 1):
        blob = rng.integers(0, 8, size).astype(np.uint8).tobytes()
        z = model_lib.pack(blob, 1)
        assert gzip.decompress(z) == blob
        e, out = model_lib.unpack_indexed(z, size)
        assert e == 0 and out == blob


def test_corruption_is_detected():
    blob = np.random.default_rng(1).standard_normal(50000, dtype=np.float32).tobytes()
    z = bytearray(model_lib.pack(blob, 1))
    z[len(z) // 2] ^= 0x10
    e1, _ = model_lib.unpack_indexed(bytes(z), len(blob))
    e2, _ = model_lib.gunzip(bytes(z), len(blob))
    assert e1 != 0 and e2 != 0
    with pytest.raises(Exception):
        gzip.decompress(bytes(z))
    z = bytearray(model_lib.pack(blob, 1))
    z[-6] ^= 1                             # CRC field
    assert model_lib.unpack_indexed(bytes(z), len(blob))[0] == 4
    assert model_lib.gunzip(bytes(z), len(blob))[0] == 4
    for cut in (5, 40, len(z) // 2, len(z) - 3):
        assert model_lib.gunzip(bytes(z[:cut]), len(blob))[0] != 0
    assert model_lib.gunzip(os.urandom(64), 1000)[0] != 0


def test_block_parallel_foreign_decoder_model():
    """mf_inflate_par.h (scan -> size -> walk -> decode to 16-bit symbols -> resolve -> verify) on zlib streams of every
    level, multi-member files, fixed-Huffman and stored blocks, header fields; same bytes as zlib, errors detected."""
    import io
    rng = np.random.default_rng(9)
    text = open(gzip.__file__, "rb").read() * 20
    cases = dict(_cases())
    cases["text_long"] = text
    cases["normal_big"] = rng.standard_normal(400000, dtype=np.float32).tobytes()
    cases["lowent_big"] = rng.integers(0, 16, 300000).astype(np.float32).tobytes()
    multi_block = 0
    for name, blob in cases.items():
        for level in (0, 1, 3, 6, 9):
            z = gzip.compress(blob, level)
            e, out, (ncand, nblocks, nmark) = model_lib.gunzip_parallel(z, len(blob))
            assert e == 0 and out == blob, (name, level, e)
            multi_block += nblocks > 1
            if level >= 3 and name == "lowent_big":
                assert nblocks >= 2 and nmark > 0, (name, level, nblocks, nmark)     # matches do reach across block boundaries
    assert multi_block > 10
    # multi-member + zero padding, FNAME/FEXTRA header fields, fixed Huffman
    a, b = cases["text"], cases["normal"]
    z = gzip.compress(a, 9) + gzip.compress(b, 1) + b"\0" * 7 + gzip.compress(b"", 3)
    e, out, _ = model_lib.gunzip_parallel(z, len(a) + len(b))
    assert e == 0 and out == a + b
    buf = io.BytesIO()
    with gzip.GzipFile(filename="named.bin", fileobj=buf, mode="wb", compresslevel=3) as f:
        f.write(b)
    e, out, _ = model_lib.gunzip_parallel(buf.getvalue(), len(b))
    assert e == 0 and out == b
    co = zlib.compressobj(9, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)
    data = b"abcabcabc" * 500 + bytes(3000)
    z = co.compress(data) + co.flush()
    e, out, (_, nblocks, _) = model_lib.gunzip_parallel(z, len(data))
    assert e == 0 and out == data
    # our own files are ordinary gzip too
    ours = model_lib.pack(cases["normal_big"], 1)
    e, out, _ = model_lib.gunzip_parallel(ours, len(cases["normal_big"]))
    assert e == 0 and out == cases["normal_big"]
    # corruption: CRC, ISIZE, payload, truncation, undersized destination
    z = bytearray(gzip.compress(cases["normal_big"], 3))
    z[-6] ^= 1
    assert model_lib.gunzip_parallel(bytes(z), len(cases["normal_big"]))[0] == 4
    z[-6] ^= 1; z[-1] ^= 1
    assert model_lib.gunzip_parallel(bytes(z), len(cases["normal_big"]))[0] == 5
    z[-1] ^= 1; z[len(z) // 2] ^= 0x40
    assert model_lib.gunzip_parallel(bytes(z), len(cases["normal_big"]))[0] != 0
    z[len(z) // 2] ^= 0x40
    assert model_lib.gunzip_parallel(bytes(z[: len(z) // 2]), len(cases["normal_big"]))[0] != 0
    e, _, _ = model_lib.gunzip_parallel(bytes(z), len(cases["normal_big"]) - 1)
    assert e == 1


def test_fast_header_parser_equals_the_portable_one():
    """mf_read_header_fast_t (7-bit table for the code-length code, word reader) either declines (0) or returns exactly what
    mf_read_dynamic_tables returns: on zlib's headers, on ours (where the bit count must equal the index entry), and on
    mutated / random headers."""
    rng = np.random.default_rng(77)

    def check(buf, comp=None, want_bits=None):
        fast, port = model_lib.header_both(buf, comp)
        if fast is None:
            assert want_bits is None, "the quick parser declined a good header"
            return False
        assert port[0] == 0, (fast[:3], port[:4])
        assert fast[0] == port[1] and fast[1] == port[2] and fast[2] == port[3] and fast[3] == port[4]
        if want_bits is not None:
            assert fast[0] == want_bits
        return True

    samples = [rng.standard_normal(50000, dtype=np.float32).tobytes(), rng.integers(0, 4, 200000, dtype=np.uint8).tobytes(),
               (b"the quick brown fox jumps over the lazy dog %d " * 4000) % tuple(range(4000)), bytes(rng.integers(0, 256, 150000, dtype=np.uint8)),
               np.arange(60000, dtype=np.int32).tobytes(), rng.integers(0, 256, 100000, dtype=np.uint8).repeat(3).tobytes()]
    headers, accepted = [], 0
    for data in samples:
        for level in (1, 3, 6, 9):
            co = zlib.compressobj(level, zlib.DEFLATED, -15)
            z = co.compress(data) + co.flush()
            if (z[0] & 7) == 4:                                  # BFINAL 0, BTYPE 10
                accepted += check(z)
                headers.append(z[:400])
        z = model_lib.pack(data)
        raw_len, nc = int.from_bytes(z[20:24], "little"), int.from_bytes(z[24:28], "little")
        off = 32 + 72 * nc
        for c in range(nc):
            comp = int.from_bytes(z[32 + 8 * c:36 + 8 * c], "little")
            bits = int.from_bytes(z[36 + 8 * c:38 + 8 * c], "little")
            if bits != 0xFFFF:                                   # not a stored chunk
                accepted += check(z[off:], comp, bits)
                headers.append(z[off:off + 400])
            off += comp
    assert accepted >= 40 and len(headers) >= 40
    # a header cut short: the chunk ends inside it
    for h in headers[:8]:
        fast, port = model_lib.header_both(h, 20)
        assert fast is None and port[0] != 0
    # mutations: single bit flips in the first 100 bytes, and random tails behind a valid first byte
    flips = 0
    for h in headers:
        for _ in range(60):
            m = bytearray(h)
            bit = int(rng.integers(0, 800))
            m[bit >> 3] ^= 1 << (bit & 7)
            flips += check(bytes(m))
    for _ in range(3000):
        m = bytearray(rng.integers(0, 256, 300, dtype=np.uint8).tobytes())
        m[0] = (m[0] & 0xF8) | 4
        check(bytes(m))
    assert flips > 100          # many mutated headers are still well-formed: the comparison above really ran


def test_decoder_input_ring_arithmetic_for_every_lane():
    """MfRingReader's address arithmetic (mf_inflate_core.h, the expressions the kernel runs): for every lane and ring
    position the read pointer walks quad after quad through the four slots and never leaves its lane's 16 bytes; with the
    refill policy of the kernel (fill every free slot at a refill point, at most two groups in flight, <= 16 bytes read
    between points) every word read belongs to a quad that has landed and no slot is overwritten before it was read."""
    import random
    L = model_lib.lib()
    R = 2048
    for base in (0x9800, 0x9800 + 7 * R, 0x1009800):              # first / last warp's ring; a non-zero CTA window in the high bits
        for lane in range(32):
            ring_lane = base + lane * 16
            # plain walk
            a, seq = ring_lane, []
            for _ in range(64):
                seq.append(a)
                a = L.model_ring_next_word(a)
            assert seq == [base + ((i // 4) % 4) * 512 + lane * 16 + (i % 4) * 4 for i in range(64)], lane
            w = ring_lane
            for k in range(9):
                assert w == base + (k % 4) * 512 + lane * 16
                w = L.model_ring_next_slot(w)
    # the kernel's policy, driven through the same expressions
    for lane in (0, 5, 31):
        for skip_words in range(4):
            for seed in range(6):
                rnd = random.Random(seed * 131 + lane)
                ring_lane = 0xD000 + lane * 16
                wslot, raddr = ring_lane, ring_lane + 4 * skip_words
                slots, groups, cur, landed, state = {}, [], [], set(), {"qn": 0, "reading": 0}

                def issue():
                    nonlocal wslot
                    old = slots.get(wslot)
                    assert old is None or old < state["reading"], "slot overwritten before it was read"
                    slots[wslot] = state["qn"]; cur.append(state["qn"]); state["qn"] += 1
                    wslot = L.model_ring_next_slot(wslot)

                def commit_and_wait(n):
                    groups.append(list(cur)); cur.clear()
                    for g in groups[:len(groups) - n] if n else groups:
                        landed.update(g)

                def word():
                    nonlocal raddr
                    q = slots[raddr & ~15]
                    assert q in landed and q == state["reading"]
                    nxt = L.model_ring_next_word(raddr)
                    if (raddr & 15) == 12:
                        state["reading"] += 1
                    raddr = nxt
                    return q

                for _ in range(4):
                    issue()
                commit_and_wait(0)
                for _ in range(3):
                    word()
                for _ in range(300):
                    if L.model_ring_free_slots(raddr, wslot):
                        issue()
                        if L.model_ring_free_slots(raddr, wslot):
                            issue()
                    commit_and_wait(2)
                    for _ in range(rnd.randint(0, 4)):
                        word()
