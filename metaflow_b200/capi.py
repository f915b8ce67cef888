"""ctypes binding of libmfcas.so (include/mfcas.h) -- the only door from Python to the GPU codec.

Loading fails LOUDLY (GpuCasUnavailable) when the in-tree library is missing; creating a
context fails loudly when no CUDA device is visible.  There is no CPU fallback in the product.
"""
import ctypes
import os

from .exception import GpuCasUnavailable

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmfcas.so")

OK, E_DST_TOO_SMALL, E_CUDA, E_BAD_STREAM, E_BAD_CRC, E_BAD_ISIZE, E_ARG, E_NOMEM, E_TRUNC = range(9)
FLAG_NO_LZ = 1

STATUS_TEXT = {
    OK: "ok",
    E_DST_TOO_SMALL: "destination buffer too small",
    E_CUDA: "CUDA error",
    E_BAD_STREAM: "Not a gzipped file / invalid deflate stream",
    E_BAD_CRC: "CRC check failed",
    E_BAD_ISIZE: "Incorrect length of data produced",
    E_ARG: "bad argument",
    E_NOMEM: "out of memory",
    E_TRUNC: "Compressed file ended before the end-of-stream marker was reached",
}

_u64p = ctypes.POINTER(ctypes.c_uint64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_vpp = ctypes.POINTER(ctypes.c_void_p)

# every symbol include/mfcas.h declares: (name, restype, argtypes)
SIGNATURES = [
    ("mfcas_device_count", ctypes.c_int, []),
    ("mfcas_create", ctypes.c_void_p, [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32]),
    ("mfcas_destroy", None, [ctypes.c_void_p]),
    ("mfcas_last_error", ctypes.c_char_p, [ctypes.c_void_p]),
    ("mfcas_bind_thread_to_device", ctypes.c_int, [ctypes.c_int]),
    ("mfcas_alloc_pinned", ctypes.c_void_p, [ctypes.c_uint64]),
    ("mfcas_free_pinned", None, [ctypes.c_void_p]),
    ("mfcas_pack_bound", ctypes.c_uint64, [ctypes.c_uint64]),
    ("mfcas_pack_batch", ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, _vpp, _u64p, _vpp, _u64p, _u64p, ctypes.c_void_p, _i32p]),
    ("mfcas_pack_extent_count", ctypes.c_uint64, [ctypes.c_uint64]),
    ("mfcas_pack_extent_bound", ctypes.c_uint64, [ctypes.c_uint64]),
    ("mfcas_pack_batch_extents", ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, _vpp, _u64p, _vpp, _u64p, _u64p, _u64p, ctypes.c_void_p, _i32p]),
    ("mfcas_hash_batch", ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, _vpp, _u64p, ctypes.c_void_p, _i32p]),
    ("mfcas_unpacked_size", ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, _u64p, ctypes.POINTER(ctypes.c_int)]),
    ("mfcas_unpack_batch", ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, _vpp, _u64p, _vpp, _u64p, _u64p, _i32p]),
    ("mfcas_pack_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, _vpp, _u64p, _vpp, _u64p, _u64p, ctypes.c_void_p, _i32p]),
    ("mfcas_pack_device_extents", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, _vpp, _u64p, _vpp, _u64p, _u64p, _u64p, ctypes.c_void_p, _i32p]),
    ("mfcas_compact_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, _vpp, _u64p, _u64p, _u64p]),
    ("mfcas_unpack_device", ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, _vpp, _u64p, _vpp, _u64p, _u64p, _i32p]),
    ("mfcas_stream", ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    ("mfcas_counters", None, [ctypes.c_void_p, _u64p, _u64p, _u64p]),
    ("mfcas_last_kernel_ms", None, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]),
    ("mfcas_version", ctypes.c_char_p, []),
]

_lib = None


def load():
    """Load libmfcas.so (built by `make -C metaflow_b200/csrc` or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpuCasUnavailable(
            "%s is missing: build it with `make -C metaflow_b200/csrc` (nvcc, sm_100a). "
            "The artifact hot path has no CPU fallback." % LIB_PATH
        )
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # see mfcas_create: one work queue per wave stream
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise GpuCasUnavailable("cannot load %s: %s" % (LIB_PATH, e))
    for name, res, args in SIGNATURES:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
