"""Pins the oracle: FIPS / RFC known answers + the fixtures the REAL reference generated
(tests/golden/cas_golden.json, made by oracle/make_golden.py importing /root/reference)."""
import base64
import gzip
import hashlib
import pickle
import zlib

import numpy as np
import pytest

import golden_inputs
from oracle import c_oracle, ref_path


def test_sha1_known_answers():
    # FIPS 180-4 / RFC 3174 vectors
    kat = {
        b"": "da39a3ee5e6b4b0d3255bfef95601890afd80709",
        b"abc": "a9993e364706816aba3e25717850c26c9cd0d89d",
        b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq": "84983e441c3bd26ebaae4aa1f95129e5e54670f1",
        b"a" * 1000000: "34aa973cd4c4daa4f61eeb2bdbad27316534016f",
    }
    for msg, hexd in kat.items():
        assert c_oracle.sha1(msg).hex() == hexd
        assert ref_path.blob_key(msg) == hexd


def test_crc32_known_answers():
    assert c_oracle.crc32(b"123456789") == 0xCBF43926
    assert c_oracle.crc32(b"") == 0
    data = bytes(range(256)) * 33
    assert c_oracle.crc32(data) == zlib.crc32(data) == ref_path.crc32(data)
    assert c_oracle.crc32(data[100:], c_oracle.crc32(data[:100])) == zlib.crc32(data)


def test_oracle_against_reference_raw_blobs(golden):
    blobs = golden_inputs.raw_blobs()
    assert set(blobs) == set(golden["raw_blobs"])
    for name, blob in blobs.items():
        g = golden["raw_blobs"][name]
        assert ref_path.blob_key(blob) == g["key"] == c_oracle.sha1(blob).hex()
        assert g["meta"] == ref_path.CAS_META_V1
        gz = base64.b64decode(g["gz_b64"])
        # reference-written file: header shape of gzip.py:259-285 (FLG 0, OS 255), decodes to the blob
        assert gz[:4] == b"\x1f\x8b\x08\x00" and gz[9] == 255
        assert ref_path.unpack_v1(gz) == blob
        e, out = c_oracle.gunzip(gz, len(blob))
        assert e == 0 and out == blob
        # the restated writer differs from the reference's file only in MTIME (bytes 4..8)
        mine = ref_path.pack_v1(blob, mtime=int.from_bytes(gz[4:8], "little"))
        assert mine == gz


def test_oracle_against_reference_artifacts(golden):
    objs = golden_inputs.artifact_objects()
    dj = golden["data_json"]
    assert dj["datastore"] == "local" and dj["version"] == "1.0" and dj["attempt"] == 0
    for name, obj in objs.items():
        blob, info = ref_path.pickle_artifact(obj)
        assert dj["info"][name] == info
        assert dj["objects"][name] == ref_path.blob_key(blob) == golden["artifacts"][name]["key"]
        gz = base64.b64decode(golden["artifacts"][name]["gz_b64"])
        e, out = c_oracle.gunzip(gz, len(blob))
        assert e == 0 and out == blob
        back = pickle.loads(out)
        if isinstance(obj, np.ndarray):
            assert back.dtype == obj.dtype and np.array_equal(back, obj)
        else:
            assert back == obj
    assert base64.b64decode(golden["done_lock_b64"]) == b'""'
    assert dj["objects"]["true"] == dj["objects"]["dup_of_true"]  # identical values share one key


@pytest.mark.parametrize("level", [0, 1, 3, 6, 9])
def test_c_inflate_matches_zlib(level):
    rng = np.random.default_rng(level)
    cases = [b"", b"x", rng.bytes(70000), b"hello world " * 5000,
             rng.standard_normal(50000, dtype=np.float32).tobytes(), bytes(100000)]
    for data in cases:
        z = gzip.compress(data, level)
        e, out = c_oracle.gunzip(z, len(data))
        assert e == 0 and out == data
        e, out = c_oracle.gunzip(z + z + b"\0" * 7, 2 * len(data))  # multi-member + zero padding
        assert e == 0 and out == data + data


def test_c_oracle_rejects_corruption():
    z = bytearray(gzip.compress(b"hello" * 1000, 3))
    z[-5] ^= 1
    assert c_oracle.gunzip(bytes(z), 5000)[0] == -4  # CRC
    z = bytearray(gzip.compress(b"hello" * 1000, 3))
    z[-1] ^= 1
    assert c_oracle.gunzip(bytes(z), 5000)[0] == -5  # ISIZE
    assert c_oracle.gunzip(b"not gzip at all", 100)[0] == -2
    assert c_oracle.gunzip(gzip.compress(b"abc" * 100)[:-9], 300)[0] == -1


def test_stored_writer_is_valid_gzip():
    for data in (b"", b"a", bytes(range(256)) * 1000):
        assert gzip.decompress(c_oracle.gzip_stored(data, mtime=5)) == data


def test_fs_layout_of_restated_path(tmp_path, golden):
    blobs = golden_inputs.raw_blobs()
    names = list(blobs)
    res = ref_path.save_blobs(str(tmp_path), "GoldenFlow/data", [blobs[n] for n in names])
    assert [k for _, k in res] == [golden["raw_blobs"][n]["key"] for n in names]
    back = dict(ref_path.load_blobs(str(tmp_path), "GoldenFlow/data", [k for _, k in res]))
    for n, (_, k) in zip(names, res):
        assert back[k] == blobs[n]
        assert hashlib.sha1(blobs[n]).hexdigest() == k
