"""The C-ABI library loads on a machine without a GPU and exports every symbol include/mfcas.h
declares; host-only entry points work; GPU entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import model_lib
from metaflow_b200 import capi
from metaflow_b200.exception import GpuCasUnavailable

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported_and_bound():
    with open(os.path.join(REPO, "include", "mfcas.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(mfcas_[a-z_0-9]+)\s*\(", header))
    declared.discard("mfcas_ctx")
    bound = {name for name, _, _ in capi.SIGNATURES}
    assert declared == bound, (declared ^ bound)
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.mfcas_version()


def test_pack_bound_matches_format_constants():
    lib = capi.load()
    assert lib.mfcas_pack_bound(0) == 45
    assert lib.mfcas_pack_bound(1) == 45 + 72 + 5 + 1
    n = 16777379                                   # pickle of a 16 MiB fp32 array
    chunks = -(-n // 32768)
    members = -(-n // (64 * 32768))                # a member is one 2 MiB stripe (64 chunks)
    assert lib.mfcas_pack_bound(n) == members * 45 + chunks * 77 + n
    assert lib.mfcas_pack_bound(n) == model_lib.lib().model_pack_bound(n)


def test_unpacked_size_is_host_only():
    lib = capi.load()
    import gzip
    import numpy as np

    blob = np.random.default_rng(0).standard_normal(30000, dtype=np.float32).tobytes()
    for packed, want_exact in ((model_lib.pack(blob), 1), (gzip.compress(blob, 3), 0)):
        size, exact = ctypes.c_uint64(0), ctypes.c_int(-1)
        rc = lib.mfcas_unpacked_size(packed, len(packed), ctypes.byref(size), ctypes.byref(exact))
        assert rc == capi.OK and size.value == len(blob) and exact.value == want_exact
    size = ctypes.c_uint64(0)
    assert lib.mfcas_unpacked_size(b"garbage-not-gzip-at-all", 23, ctypes.byref(size), None) == capi.E_BAD_STREAM


@pytest.mark.skipif(capi.load().mfcas_device_count() > 0, reason="only meaningful without a GPU")
def test_no_gpu_means_loud_failure():
    lib = capi.load()
    assert lib.mfcas_create(None, 0, 0, 0) is None
    from metaflow_b200.codec import GpuCodec

    with pytest.raises(GpuCasUnavailable):
        GpuCodec()
    from metaflow_b200 import FlowDataStore, LocalStorage

    fds = FlowDataStore("F", storage_impl=LocalStorage, ds_root="/tmp/_mfcas_nogpu")
    with pytest.raises(GpuCasUnavailable):
        fds.ca_store.save_blobs(iter([b"abc"]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "metaflow_b200")
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(root, fn), errors="replace") as f:
                    text = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), os.path.join(root, fn)
                natives = r"zlib|gzip" if fn == "filecache.py" else r"zlib|gzip|hashlib"
                # (client/filecache.py names its metadata cache files sha1("run/step/task/attempt/metadata"), as the
                # reference does, filecache.py:431-439 -- a file NAME, never artifact bytes)
                assert not re.search(r"^\s*(from|import)\s+(%s)\b" % natives, text, re.M), os.path.join(root, fn)


def test_codec_slot_layout_is_vectorised_and_aligned():
    """GpuCodec._slots: 256-byte aligned slots with a 256-byte gap -- the layout the page-locked arenas use."""
    import numpy as np
    from metaflow_b200.codec import GpuCodec
    sizes = np.array([0, 1, 255, 256, 257, 65697, 16777379], dtype=np.uint64)
    offs, total = GpuCodec._slots(sizes)
    assert offs.dtype == np.uint64 and offs[0] == 0
    for i in range(len(sizes)):
        assert int(offs[i]) % 256 == 0
        end = int(offs[i]) + int(sizes[i])
        nxt = int(offs[i + 1]) if i + 1 < len(sizes) else total
        assert nxt - end >= 256 and nxt - end < 512 + 256
    assert GpuCodec._slots(np.zeros(0, dtype=np.uint64))[1] == 0


def test_file_source_reads_into_a_buffer(tmp_path):
    import numpy as np
    from metaflow_b200.codec import FileSource
    data = bytes(range(256)) * 300
    p = tmp_path / "blob"
    p.write_bytes(data)
    fs = FileSource(str(p), len(data))
    dst = np.full(len(data) + 100, 0xAA, dtype=np.uint8)
    fs.read_into(dst, 50)
    assert dst[50:50 + len(data)].tobytes() == data and (dst[:50] == 0xAA).all() and (dst[50 + len(data):] == 0xAA).all()
    assert len(fs) == len(data) and fs.head(10) == data[:10]
    short = FileSource(str(p), len(data) + 5)          # the file is shorter than promised
    import pytest
    from metaflow_b200.exception import DataException
    with pytest.raises(DataException):
        short.read_into(np.zeros(len(data) + 5, dtype=np.uint8), 0)
