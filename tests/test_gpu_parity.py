"""GPU parity tests proper (run with -m gpu on a B200): the CUDA path, called through the C ABI
(GpuCodec -> libmfcas.so), against the CPU oracle on the same inputs.  Bit-exact everywhere:
keys == hashlib.sha1, decoded bytes == input, our files decode with CPython gzip (what the
reference's _unpack_v1 runs) and with the independent C oracle, reference-written files decode
on the GPU."""
import base64
import ctypes
import gzip
import hashlib
import json
import os
import pickle
import zlib

import numpy as np
import pytest

import golden_inputs
from oracle import c_oracle, ref_path

pytestmark = pytest.mark.gpu


def _roundtrip(codec, blobs):
    keys, packed = codec.pack_batch(blobs)
    packed = [bytes(p) for p in packed]
    for i, (b, k, z) in enumerate(zip(blobs, keys, packed)):
        b = bytes(b)
        assert k == hashlib.sha1(b).hexdigest(), i
        assert z[:4] == b"\x1f\x8b\x08\x04", i
        assert gzip.decompress(z) == b, i
        e, out = c_oracle.gunzip(z, len(b))
        assert e == 0 and out == b, (i, e)
    back = codec.unpack_batch(packed)
    for i, (b, o) in enumerate(zip(blobs, back)):
        assert not isinstance(o, Exception), (i, o)
        assert o == bytes(b), i
    return keys, packed


def test_golden_raw_blobs(gpu_codec, golden):
    blobs = golden_inputs.raw_blobs()
    names = list(blobs)
    keys, _ = _roundtrip(gpu_codec, [blobs[n] for n in names])
    assert keys == [golden["raw_blobs"][n]["key"] for n in names]
    assert gpu_codec.hash_batch([blobs[n] for n in names]) == keys


def test_golden_artifacts_and_reference_written_files(gpu_codec, golden):
    objs = golden_inputs.artifact_objects()
    names = list(objs)
    pickles = [pickle.dumps(objs[n], protocol=4) for n in names]
    keys, _ = _roundtrip(gpu_codec, pickles)
    assert keys == [golden["data_json"]["objects"][n] for n in names]
    # files the REAL reference wrote (zlib level 3, single stream) decode on the GPU
    ref_files = [base64.b64decode(golden["artifacts"][n]["gz_b64"]) for n in names]
    raw = golden_inputs.raw_blobs()
    ref_files += [base64.b64decode(golden["raw_blobs"][n]["gz_b64"]) for n in raw]
    want = pickles + [raw[n] for n in raw]
    got = gpu_codec.unpack_batch(ref_files)
    for i, (w, g) in enumerate(zip(want, got)):
        assert not isinstance(g, Exception), (i, g)
        assert g == w, i


def test_edge_sizes_ragged_batch(gpu_codec):
    rng = np.random.default_rng(7)
    sizes = [0, 1, 2, 3, 4, 5, 15, 16, 17, 55, 56, 57, 63, 64, 65, 119, 120, 127, 128, 1023, 1024, 1025, 2047, 2048, 2049,
             4095, 32767, 32768, 32769, 65535, 65536, 65537, 100000, 0, 1,
             2097151, 2097152, 2097153, 4194304 + 77]      # stripe (2 MiB) boundaries of the striped packer
    blobs = []
    for i, n in enumerate(sizes):
        kind = i % 4
        if kind == 0:
            blobs.append(rng.bytes(n))
        elif kind == 1:
            blobs.append(bytes(n))
        elif kind == 2:
            blobs.append((b"metaflow foreach join artifact " * (n // 31 + 1))[:n])
        else:
            blobs.append(rng.integers(0, 5, n).astype(np.uint8).tobytes())
    _roundtrip(gpu_codec, blobs)


def test_striped_and_contiguous_packers_agree(gpu_codec):
    """mfcas_pack_batch_extents (2 MiB stripes written straight into extents) and mfcas_pack_batch (the same members,
    moved together by the compaction kernels) give the same keys and the same bytes; a blob is a multi-member gzip file."""
    rng = np.random.default_rng(17)
    blobs = [rng.standard_normal(n, dtype=np.float32).tobytes() for n in (10, 300000, 1500000, 2500000)] + [b"", bytes(5 << 20)]
    k1, p1 = gpu_codec.pack_batch(blobs)
    p1 = [bytes(p) for p in p1]
    k2, p2 = gpu_codec.pack_batch_contiguous(blobs)
    p2 = [bytes(p) for p in p2]
    assert k1 == k2 == [hashlib.sha1(b).hexdigest() for b in blobs]
    for b, z1, z2 in zip(blobs, p1, p2):
        assert gzip.decompress(z1) == b and gzip.decompress(z2) == b
        e, out = c_oracle.gunzip(z1, len(b))
        assert e == 0 and out == b
    assert gpu_codec.unpack_batch(p1) == blobs and gpu_codec.unpack_batch(p2) == blobs
    assert p1 == p2                           # same members, whichever way they reach the caller
    assert p1[3].count(b"\x1f\x8b\x08\x04") >= 5    # 10 MB of float32 = 5 stripes = 5 gzip members


def test_distributions_and_ratio(gpu_codec):
    rng = np.random.default_rng(11)
    n = 1 << 18
    data = {
        "normal": rng.standard_normal(n, dtype=np.float32).tobytes(),
        "uniform": rng.random(n, dtype=np.float32).tobytes(),
        "lowent": rng.integers(0, 16, n).astype(np.float32).tobytes(),
        "zeros": bytes(4 * n),
        "text": (open(gzip.__file__, "rb").read() * 40)[: 4 * n],
        "int64": np.arange(n // 2, dtype=np.int64).tobytes(),
    }
    _, packed = _roundtrip(gpu_codec, list(data.values()))
    ratios = {k: len(z) / len(b) for (k, b), z in zip(data.items(), packed)}
    print("ratios", {k: round(v, 4) for k, v in ratios.items()})
    for k in ("normal", "uniform"):
        assert ratios[k] <= len(gzip.compress(data[k], 1)) / len(data[k]) * 1.005, k   # >= zlib level 1 on float noise
    assert ratios["zeros"] < 0.02 and ratios["lowent"] < 0.32 and ratios["text"] < 0.6


def test_codes_carried_inside_a_group(gpu_codec):
    """The encoder keeps the Huffman codes of the previous chunk of a 64-chunk group while they stay within 1/128 of
    fresh ones.  Mix chunk kinds inside groups (stationary noise, sudden switches, symbols the kept code does not
    cover, incompressible and empty-ish chunks): every stream must stay valid for zlib and for our decoder, packed
    bytes must not depend on the batch around a blob, and the kept codes must not cost ratio on stationary data."""
    rng = np.random.default_rng(23)
    K = 32768
    noise = lambda n: rng.standard_normal(n // 4, dtype=np.float32).tobytes()
    text = (open(gzip.__file__, "rb").read() * 8)
    kinds = [
        noise(40 * K),                                                           # stationary: codes reused
        b"".join(noise(K) if i % 2 else text[i * K:(i + 1) * K] for i in range(40)),   # alternating every chunk
        noise(5 * K) + bytes(3 * K) + noise(5 * K) + bytes(range(256)) * 512 + noise(3 * K + 777),
        b"".join(rng.integers(0, 1 << (1 + i % 8), K, dtype=np.uint8).tobytes() for i in range(37)),   # alphabet grows and shrinks
        rng.integers(0, 256, 20 * K + 5, dtype=np.uint8).tobytes(),             # incompressible: stored chunks
        bytes(17 * K) + b"x",
    ]
    _, packed = _roundtrip(gpu_codec, kinds)
    # a blob's packed bytes are a function of the blob alone
    _, alone = gpu_codec.pack_batch([kinds[1]])
    assert bytes(alone[0]) == packed[1]
    _, again = gpu_codec.pack_batch(list(reversed(kinds)))
    assert [bytes(p) for p in reversed(again)] == packed
    ratio = len(packed[0]) / len(kinds[0])
    assert ratio <= len(gzip.compress(kinds[0], 1)) / len(kinds[0]) * 1.005, ratio


def test_member_boundaries(gpu_codec):
    rng = np.random.default_rng(3)
    n = 768 * 32768
    blobs = [rng.integers(0, 8, size).astype(np.uint8).tobytes() for size in (n - 1, n, n + 1, 2 * n + 12345)]
    _roundtrip(gpu_codec, blobs)


def test_foreign_streams(gpu_codec):
    rng = np.random.default_rng(5)
    payloads = [b"", b"x", rng.bytes(70000), b"hello world " * 5000, rng.standard_normal(80000, dtype=np.float32).tobytes(), bytes(300000)]
    files, want = [], []
    for p in payloads:
        for level in (0, 1, 3, 9):
            files.append(gzip.compress(p, level)); want.append(p)
    files.append(gzip.compress(payloads[2], 3) + gzip.compress(payloads[3], 9) + b"\0" * 9); want.append(payloads[2] + payloads[3])
    co = zlib.compressobj(9, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)
    files.append(co.compress(payloads[3]) + co.flush()); want.append(payloads[3])
    import io
    buf = io.BytesIO()
    with gzip.GzipFile(filename="named.bin", fileobj=buf, mode="wb", compresslevel=3) as f:   # FNAME header field
        f.write(payloads[4])
    files.append(buf.getvalue()); want.append(payloads[4])
    files.append(c_oracle.gzip_stored(payloads[2], mtime=3)); want.append(payloads[2])
    got = gpu_codec.unpack_batch(files)
    for i, (w, g) in enumerate(zip(want, got)):
        assert not isinstance(g, Exception), (i, g)
        assert g == w, i


def test_corruption_and_error_text(gpu_codec, tmp_path):
    from metaflow_b200 import capi
    blob = np.random.default_rng(1).standard_normal(60000, dtype=np.float32).tobytes()
    _, packed = gpu_codec.pack_batch([blob])
    good = bytes(packed[0])
    crc_flip = bytearray(good); crc_flip[-6] ^= 1
    isize_flip = bytearray(good); isize_flip[-1] ^= 1
    payload_flip = bytearray(good); payload_flip[len(good) // 2] ^= 0x20
    ref = bytearray(gzip.compress(blob, 3)); ref[-6] ^= 1
    ref_trunc = gzip.compress(blob, 3)[:-20]
    cases = [bytes(crc_flip), bytes(isize_flip), bytes(payload_flip), bytes(ref), ref_trunc, b"definitely not gzip data", good[: len(good) // 2], good]
    out = gpu_codec.unpack_batch(cases)
    assert out[0].code == capi.E_BAD_CRC and "CRC check failed" in str(out[0])
    assert isinstance(out[1], Exception) and isinstance(out[2], Exception)
    assert out[3].code == capi.E_BAD_CRC
    assert isinstance(out[4], Exception) and isinstance(out[5], Exception) and isinstance(out[6], Exception)
    assert out[7] == blob                                  # a good blob in the same batch is unaffected
    # through the datastore class the failure surfaces as the reference's DataException text
    from metaflow_b200 import FlowDataStore, LocalStorage
    from metaflow_b200.exception import DataException
    fds = FlowDataStore("C", storage_impl=LocalStorage, ds_root=str(tmp_path), codec=gpu_codec)
    key = fds.ca_store.save_blobs(iter([blob]))[0].key
    p = tmp_path / "C" / "data" / key[:2] / key
    p.write_bytes(bytes(crc_flip))
    with pytest.raises(DataException) as e:
        list(fds.ca_store.load_blobs([key]))
    assert "Could not unpack artifact 'C/data/%s/%s'" % (key[:2], key) in str(e.value) and "CRC check failed" in str(e.value)


def test_datastore_end_to_end_on_gpu(gpu_codec, tmp_path, golden):
    from metaflow_b200 import FlowDataStore, LocalStorage
    objs = golden_inputs.artifact_objects()
    fds = FlowDataStore("GoldenFlow", storage_impl=LocalStorage, ds_root=str(tmp_path), codec=gpu_codec)
    w = fds.get_task_datastore("1", "start", "7", attempt=0, mode="w")
    w.init_task(); w.save_artifacts(iter(objs.items()), len_hint=len(objs)); w.done()
    dj = json.loads((tmp_path / "GoldenFlow" / "1" / "start" / "7" / "0.data.json").read_text())
    assert dj["objects"] == golden["data_json"]["objects"] and dj["info"] == golden["data_json"]["info"]
    for name, key in dj["objects"].items():                # every stored file is plain gzip to the reference
        z = (tmp_path / "GoldenFlow" / "data" / key[:2] / key).read_bytes()
        assert ref_path.unpack_v1(z) == pickle.dumps(objs[name], protocol=4)
        assert json.loads((tmp_path / "GoldenFlow" / "data" / key[:2] / (key + "_meta")).read_text()) == {"cas_raw": False, "cas_version": 1}
    back = dict(fds.get_task_datastore("1", "start", "7").load_artifacts(list(objs)))
    for n, o in objs.items():
        assert (np.array_equal(o, back[n]) and o.dtype == back[n].dtype) if isinstance(o, np.ndarray) else o == back[n], n
    # files written by the reference's code path (restated: gzip level 3) load through the GPU
    ref_root = tmp_path / "ref"
    objects, info = ref_path.save_artifacts(str(ref_root), "GoldenFlow", objs.items())
    assert objects == dj["objects"] and info == dj["info"]
    fds2 = FlowDataStore("GoldenFlow", storage_impl=LocalStorage, ds_root=str(ref_root), codec=gpu_codec)
    got = dict(fds2.ca_store.load_blobs(list(set(objects.values()))))
    for name, key in objects.items():
        assert got[key] == pickle.dumps(objs[name], protocol=4)


def test_no_lz_flag_and_wave_splitting():
    from metaflow_b200.codec import GpuCodec
    rng = np.random.default_rng(2)
    blobs = [rng.integers(0, 3, 300000).astype(np.uint8).tobytes() for _ in range(9)] + [b"", b"tiny"]
    c = GpuCodec(wave_bytes=1 << 20, lz=False)             # 1 MiB waves -> the batch is cut into several
    try:
        _roundtrip(c, blobs)
        assert c.counters()["kernel_launches"] >= 3 * 4
    finally:
        c.close()


def test_baseline_size_properties(gpu_codec):
    """BASELINE.json config-2 blob size (16 MiB fp32 -> 16,777,379-byte pickle), size-independent
    properties: encode -> decode round trip, keys == hashlib, checksum of checksums."""
    n = 12
    arrays = [np.random.default_rng(100 + i).standard_normal(4194304, dtype=np.float32) for i in range(n)]
    blobs = [pickle.dumps(a, protocol=4) for a in arrays]
    assert len(blobs[0]) == 16777379
    keys, packed = gpu_codec.pack_batch(blobs)
    packed = [bytes(p) for p in packed]
    assert keys == [hashlib.sha1(b).hexdigest() for b in blobs]
    back = gpu_codec.unpack_batch(packed)
    agg = hashlib.sha1()
    for b, o in zip(blobs, back):
        assert o == b
        agg.update(hashlib.sha1(o).digest())
    want = hashlib.sha1()
    for k in keys:
        want.update(bytes.fromhex(k))
    assert agg.digest() == want.digest()
    assert gzip.decompress(packed[3]) == blobs[3]           # the reference's reader on a full-size blob
    ratio = sum(map(len, packed)) / sum(map(len, blobs))
    print("16 MiB N(0,1) ratio", ratio)
    assert ratio < 0.935


def test_many_small_blobs_dense_sha_mode(gpu_codec):
    """More than 296 SHA-1 CTAs (> 9472 blobs) switches the hash kernel to its densely packed mode;
    BASELINE config C3 shape (many small, hash-bound blobs) at reduced count."""
    rng = np.random.default_rng(21)
    base = rng.standard_normal(2048 * 3, dtype=np.float32).tobytes()
    blobs = [base[(i % 97):(i % 97) + 4096 + (i % 13)] + i.to_bytes(4, "little") for i in range(12000)]
    keys, packed = gpu_codec.pack_batch(blobs)
    packed = [bytes(p) for p in packed]
    assert keys == [hashlib.sha1(b).hexdigest() for b in blobs]
    for i in (0, 1, 5999, 11999):
        assert gzip.decompress(packed[i]) == blobs[i]
    back = gpu_codec.unpack_batch(packed)
    assert all(o == b for o, b in zip(back, blobs))
    assert gpu_codec.hash_batch(blobs[:5000]) == keys[:5000]


def test_join_fan_in_is_one_batched_unpack(gpu_codec, tmp_path):
    """BASELINE config C4 shape at reduced size: a join that loads the artifacts of 64 foreach siblings gets
    them through TaskDataStoreSet's prefetch = ONE batched GPU unpack, on blobs written by us and on blobs
    written by the reference's code path (gzip level 3, no index)."""
    from metaflow_b200 import FlowDataStore, LocalStorage, TaskDataStoreSet
    from metaflow_b200.tuple_util import ForeachFrame
    rng = np.random.default_rng(41)
    arrays = [rng.standard_normal(16384, dtype=np.float32) for _ in range(64)]
    for variant in ("ours", "reference"):
        root = tmp_path / variant
        fds = FlowDataStore("JoinFlow", storage_impl=LocalStorage, ds_root=str(root), codec=gpu_codec)
        for i, a in enumerate(arrays):
            w = fds.get_task_datastore("1", "leaf", str(100 + i), attempt=0, mode="w")
            w.init_task()
            arts = [("x", a), ("_foreach_stack", [ForeachFrame("start", "items", 64, i, i)])]
            if variant == "ours":
                w.save_artifacts(iter(arts))
            else:                                   # blobs as stock Metaflow writes them
                objects, info = ref_path.save_artifacts(str(root), "JoinFlow", arts)
                w._objects.update(objects); w._info.update(info)
            w.done()
        launches0 = gpu_codec.counters()["kernel_launches"]
        dss = TaskDataStoreSet(fds, "1", steps=["leaf"], prefetch_data_artifacts=["x", "_foreach_stack"])
        launches1 = gpu_codec.counters()["kernel_launches"]
        # one batch: parse + inflate + verify + serial-foreign, plus the 7 kernels of the block-parallel foreign decoder
        # when the batch holds reference-written (index-less) streams
        assert launches1 - launches0 == (4 if variant == "ours" else 11)
        got = {ds["_foreach_stack"][0].index: ds["x"] for ds in dss}
        assert gpu_codec.counters()["kernel_launches"] == launches1   # all served from the prefetch cache
        assert all(np.array_equal(got[i], a) for i, a in enumerate(arrays))
        fds.ca_store.set_blob_cache(None)


def test_two_device_context_round_robin():
    """One context over two GPUs: blob i goes to device i mod 2 (one host thread per device), results come
    back in input order.  Skipped on a single-GPU box."""
    from metaflow_b200 import capi
    from metaflow_b200.codec import GpuCodec
    if capi.load().mfcas_device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(31)
    blobs = [rng.standard_normal(int(rng.integers(1, 300000)), dtype=np.float32).tobytes() for _ in range(41)] + [b"", b"x"]
    c = GpuCodec(devices=[0, 1], wave_bytes=8 << 20)
    try:
        _roundtrip(c, blobs)
        assert c.hash_batch(blobs) == [hashlib.sha1(b).hexdigest() for b in blobs]
    finally:
        c.close()


def test_device_resident_api(gpu_codec):
    import torch
    from metaflow_b200 import capi
    lib = capi.load()
    rng = np.random.default_rng(9)
    blobs = [rng.standard_normal(50000, dtype=np.float32).tobytes(), b"", rng.bytes(33333)]
    d_src = [torch.frombuffer(bytearray(b) + bytearray(16), dtype=torch.uint8).cuda() for b in blobs]
    caps = [lib.mfcas_pack_bound(len(b)) for b in blobs]
    d_dst = [torch.zeros(c + 16, dtype=torch.uint8, device="cuda") for c in caps]
    torch.cuda.synchronize()
    n = len(blobs)
    src = (ctypes.c_void_p * n)(*[t.data_ptr() for t in d_src]); dst = (ctypes.c_void_p * n)(*[t.data_ptr() for t in d_dst])
    slen = (ctypes.c_uint64 * n)(*[len(b) for b in blobs]); cap = (ctypes.c_uint64 * n)(*caps)
    dlen = (ctypes.c_uint64 * n)(); st = (ctypes.c_int32 * n)(); sha = ctypes.create_string_buffer(20 * n)
    assert lib.mfcas_pack_device(gpu_codec._ctx, 0, n, src, slen, dst, cap, dlen, sha, st) == 0
    assert list(st) == [0] * n
    for i, b in enumerate(blobs):
        assert sha.raw[20 * i:20 * i + 20] == hashlib.sha1(b).digest()
        z = d_dst[i][: dlen[i]].cpu().numpy().tobytes()
        assert gzip.decompress(z) == b
    d_out = [torch.zeros(len(b) + 16, dtype=torch.uint8, device="cuda") for b in blobs]
    torch.cuda.synchronize()
    out = (ctypes.c_void_p * n)(*[t.data_ptr() for t in d_out]); ocap = (ctypes.c_uint64 * n)(*[len(b) for b in blobs])
    olen = (ctypes.c_uint64 * n)(); st2 = (ctypes.c_int32 * n)()
    assert lib.mfcas_unpack_device(gpu_codec._ctx, 0, n, dst, dlen, out, ocap, olen, st2) == 0
    assert list(st2) == [0] * n
    for i, b in enumerate(blobs):
        assert olen[i] == len(b) and d_out[i][: len(b)].cpu().numpy().tobytes() == b
    assert lib.mfcas_stream(gpu_codec._ctx, 0)


def test_c_abi_contract_edges(gpu_codec):
    """The raw C ABI: undersized destinations are reported per blob (and the other blobs of the batch are untouched by
    it), empty batches and bad arguments return codes instead of crashing."""
    from metaflow_b200 import capi
    lib = capi.load()
    ctx = gpu_codec._ctx
    rng = np.random.default_rng(41)
    blobs = [rng.standard_normal(700000, dtype=np.float32).tobytes(), b"tiny", rng.integers(0, 4, 3 << 20, dtype=np.uint8).tobytes()]
    n = len(blobs)
    u64, vp = ctypes.c_uint64, ctypes.c_void_p
    src_buf = [ctypes.create_string_buffer(b, len(b)) for b in blobs]
    src = (vp * n)(*[ctypes.addressof(s) for s in src_buf]); slen = (u64 * n)(*map(len, blobs))
    st = (ctypes.c_int32 * n)(); sha = ctypes.create_string_buffer(20 * n)

    # contiguous packer: blob 0 gets a destination that is too small
    caps = [lib.mfcas_pack_bound(len(b)) for b in blobs]
    caps[0] = 1000
    dst_buf = [ctypes.create_string_buffer(c + 16) for c in caps]
    dst = (vp * n)(*[ctypes.addressof(d) for d in dst_buf]); dcap = (u64 * n)(*caps); dlen = (u64 * n)()
    assert lib.mfcas_pack_batch(ctx, n, src, slen, dst, dcap, dlen, sha, st) == capi.OK
    assert list(st) == [capi.E_DST_TOO_SMALL, 0, 0] and dlen[0] > 1000          # the need is reported
    assert sha.raw[:20] == hashlib.sha1(blobs[0]).digest()                       # the digest does not depend on the destination
    for i in (1, 2):
        assert gzip.decompress(dst_buf[i].raw[:dlen[i]]) == blobs[i]

    # striped packer: same contract per blob
    ecaps = [lib.mfcas_pack_extent_bound(len(b)) for b in blobs]
    ecaps[2] -= 1
    cnt = [lib.mfcas_pack_extent_count(len(b)) for b in blobs]
    edst_buf = [ctypes.create_string_buffer(c + 16) for c in ecaps]
    edst = (vp * n)(*[ctypes.addressof(d) for d in edst_buf]); ecap = (u64 * n)(*ecaps)
    eo = (u64 * sum(cnt))(); el = (u64 * sum(cnt))()
    assert lib.mfcas_pack_batch_extents(ctx, n, src, slen, edst, ecap, eo, el, sha, st) == capi.OK
    assert list(st) == [0, 0, capi.E_DST_TOO_SMALL]
    base = 0
    for i in range(2):
        z = b"".join(edst_buf[i].raw[eo[base + k]:eo[base + k] + el[base + k]] for k in range(cnt[i]))
        assert gzip.decompress(z) == blobs[i]
        base += cnt[i]

    # unpack into an undersized destination
    good = dst_buf[2].raw[:dlen[2]]
    psrc_buf = ctypes.create_string_buffer(good, len(good))
    psrc = (vp * 1)(ctypes.addressof(psrc_buf)); plen = (u64 * 1)(len(good))
    out = ctypes.create_string_buffer(len(blobs[2]))
    odst = (vp * 1)(ctypes.addressof(out)); ocap = (u64 * 1)(len(blobs[2]) - 1); olen = (u64 * 1)(); st1 = (ctypes.c_int32 * 1)()
    assert lib.mfcas_unpack_batch(ctx, 1, psrc, plen, odst, ocap, olen, st1) == capi.OK
    assert st1[0] == capi.E_DST_TOO_SMALL and olen[0] == len(blobs[2])
    ocap[0] = len(blobs[2])
    assert lib.mfcas_unpack_batch(ctx, 1, psrc, plen, odst, ocap, olen, st1) == capi.OK and st1[0] == 0 and out.raw == blobs[2]

    # empty batch, null arguments
    assert lib.mfcas_pack_batch(ctx, 0, None, None, None, None, None, None, None) == capi.OK
    assert lib.mfcas_unpack_batch(ctx, 0, None, None, None, None, None, None) == capi.OK
    assert lib.mfcas_pack_batch(ctx, 1, None, None, None, None, None, None, None) == capi.E_ARG
    assert lib.mfcas_pack_batch(None, 1, src, slen, dst, dcap, dlen, sha, st) == capi.E_ARG


def test_client_filecache_batched_fetch_on_gpu(gpu_codec, tmp_path):
    """the client file cache in front of the GPU datastore -- the artifacts of a whole step are
    fetched with ONE batched unpack, land in the on-disk cache, and come from there the second time."""
    from metaflow_b200 import FlowDataStore, LocalStorage
    from metaflow_b200.client import FileCache
    rng = np.random.default_rng(51)
    fds = FlowDataStore("ClientFlow", storage_impl=LocalStorage, ds_root=str(tmp_path / "ds"), codec=gpu_codec)
    arrays = [rng.standard_normal(60000 + 1000 * t, dtype=np.float32) for t in range(12)]
    for t, a in enumerate(arrays):
        w = fds.get_task_datastore("3", "train", str(t), attempt=0, mode="w")
        w.init_task()
        w.save_artifacts(iter([("weights", a), ("tag", "t%d" % t)]))
        w.done()
    fc = FileCache(cache_dir=str(tmp_path / "cache"), max_size=1000, codec=gpu_codec)
    tasks = [("3", "train", str(t), None) for t in range(12)]
    l0 = gpu_codec.counters()["kernel_launches"]
    out = fc.get_artifacts_for_tasks("local", str(tmp_path / "ds"), "ClientFlow", tasks, ["weights", "tag"])
    l1 = gpu_codec.counters()["kernel_launches"]
    assert l1 - l0 == 4                                             # parse + inflate + verify + foreign: one batch for 24 blobs
    for t, a in enumerate(arrays):
        got = out[("3", "train", str(t))]
        assert np.array_equal(got["weights"], a) and got["tag"] == "t%d" % t
    again = fc.get_artifacts_for_tasks("local", str(tmp_path / "ds"), "ClientFlow", tasks, ["weights"])
    assert gpu_codec.counters()["kernel_launches"] == l1             # all from the cache files
    assert np.array_equal(again[("3", "train", "7")]["weights"], arrays[7])
    assert np.array_equal(fc.get_artifact("local", str(tmp_path / "ds"), None, "ClientFlow", "3", "train", "5", "weights"), arrays[5])


def test_device_resident_extents_and_in_place_compaction(gpu_codec):
    """mfcas_pack_device_extents (members written straight into fixed slots of the destination) + mfcas_compact_device (moved
    together in place) = the contiguous RFC 1952 file mfcas_unpack_device takes; undersized destinations are reported per blob."""
    import torch
    from metaflow_b200 import capi
    lib = capi.load()
    rng = np.random.default_rng(19)
    blobs = [rng.standard_normal(1400000, dtype=np.float32).tobytes(), b"", rng.integers(0, 7, 2097152 + 5, dtype=np.uint8).tobytes(),
             b"short", rng.standard_normal(300, dtype=np.float32).tobytes()]
    n = len(blobs)
    u64, vp = ctypes.c_uint64, ctypes.c_void_p
    d_src = [torch.frombuffer(bytearray(b) + bytearray(16), dtype=torch.uint8).cuda() for b in blobs]
    caps = [lib.mfcas_pack_extent_bound(len(b)) for b in blobs]
    cnt = [lib.mfcas_pack_extent_count(len(b)) for b in blobs]
    caps[3] = 8                                                     # blob 3 cannot fit
    d_dst = [torch.zeros(max(c, 64) + 256, dtype=torch.uint8, device="cuda") for c in caps]
    torch.cuda.synchronize()
    src = (vp * n)(*[t.data_ptr() for t in d_src]); dst = (vp * n)(*[t.data_ptr() for t in d_dst])
    slen = (u64 * n)(*map(len, blobs)); cap = (u64 * n)(*caps)
    eo = (u64 * sum(cnt))(); el = (u64 * sum(cnt))(); st = (ctypes.c_int32 * n)(); sha = ctypes.create_string_buffer(20 * n)
    assert lib.mfcas_pack_device_extents(gpu_codec._ctx, 0, n, src, slen, dst, cap, eo, el, sha, st) == 0
    assert list(st) == [0, 0, 0, capi.E_DST_TOO_SMALL, 0]
    for i, b in enumerate(blobs):
        assert sha.raw[20 * i:20 * i + 20] == hashlib.sha1(b).digest()
    assert cnt == [3, 1, 2, 1, 1]
    dlen = (u64 * n)()
    assert lib.mfcas_compact_device(gpu_codec._ctx, 0, n, dst, slen, el, dlen) == 0
    base = 0
    for i, b in enumerate(blobs):
        if st[i] == 0:
            assert dlen[i] == sum(el[base + k] for k in range(cnt[i]))
            z = d_dst[i][: dlen[i]].cpu().numpy().tobytes()
            assert gzip.decompress(z) == b
            e, out = c_oracle.gunzip(z, len(b))
            assert e == 0 and out == b
        base += cnt[i]
    # and back through the device decoder
    good = [0, 1, 2, 4]
    m = len(good)
    d_out = [torch.zeros(len(blobs[i]) + 16, dtype=torch.uint8, device="cuda") for i in good]
    torch.cuda.synchronize()
    psrc = (vp * m)(*[d_dst[i].data_ptr() for i in good]); plen = (u64 * m)(*[dlen[i] for i in good])
    out = (vp * m)(*[t.data_ptr() for t in d_out]); ocap = (u64 * m)(*[len(blobs[i]) for i in good])
    olen = (u64 * m)(); st2 = (ctypes.c_int32 * m)()
    assert lib.mfcas_unpack_device(gpu_codec._ctx, 0, m, psrc, plen, out, ocap, olen, st2) == 0
    assert list(st2) == [0] * m
    for k, i in enumerate(good):
        assert d_out[k][: len(blobs[i])].cpu().numpy().tobytes() == blobs[i]
