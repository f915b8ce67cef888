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
    assert lib.mfcas_pack_bound(n) == 45 + chunks * 77 + n
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
                assert not re.search(r"^\s*(from|import)\s+(zlib|gzip|hashlib)\b", text, re.M), os.path.join(root, fn)
