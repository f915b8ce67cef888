"""Builds and loads tests/host_model/libmfmodel.so -- the CPU model of the MF container format."""
import ctypes
import os
import subprocess

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_model")
_SO = os.path.join(_HERE, "libmfmodel.so")
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(_HERE, "model.cpp")
    csrc = os.path.join(os.path.dirname(_HERE), "..", "metaflow_b200", "csrc")
    newest = max(os.path.getmtime(p) for p in (src, os.path.join(csrc, "mf_core.h"), os.path.join(csrc, "mf_inflate_core.h"),
                                               os.path.join(csrc, "mf_inflate_par.h")))
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _SO, src])
    L = ctypes.CDLL(_SO)
    L.model_pack.restype = ctypes.c_uint64
    L.model_pack.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
    L.model_pack_bound.restype = ctypes.c_uint64
    L.model_pack_bound.argtypes = [ctypes.c_uint64]
    sig = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    L.model_gunzip.argtypes = sig
    L.model_unpack_indexed.argtypes = sig
    L.model_gunzip_parallel.argtypes = sig + [ctypes.POINTER(ctypes.c_uint64)]
    L.model_header_both.restype = None
    L.model_header_both.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    for fn, nargs in ((L.model_ring_next_word, 1), (L.model_ring_next_slot, 1), (L.model_ring_free_slots, 2)):
        fn.restype = ctypes.c_uint32
        fn.argtypes = [ctypes.c_uint32] * nargs
    _lib = L
    return L


def pack(blob, lz=1):
    L = lib()
    cap = L.model_pack_bound(len(blob))
    out = ctypes.create_string_buffer(cap)
    n = L.model_pack(blob, len(blob), out, cap, lz)
    assert n > 0
    return out.raw[:n]


def _unp(fn, z, cap):
    out = ctypes.create_string_buffer(max(cap, 1))
    p = ctypes.c_uint64(0)
    e = fn(z, len(z), out, cap, ctypes.byref(p))
    return e, out.raw[: p.value]


def gunzip(z, cap):
    return _unp(lib().model_gunzip, z, cap)


def unpack_indexed(z, cap):
    return _unp(lib().model_unpack_indexed, z, cap)


def gunzip_parallel(z, cap):
    """-> (error code, bytes, (candidates, chain blocks, markers)) through the block-parallel foreign decoder's model."""
    out = ctypes.create_string_buffer(max(cap, 1))
    p = ctypes.c_uint64(0)
    stats = (ctypes.c_uint64 * 3)()
    e = lib().model_gunzip_parallel(z, len(z), out, cap, ctypes.byref(p), stats)
    return e, out.raw[: p.value] if e == 0 else b"", tuple(stats)


def header_both(src, comp_bytes=None):
    """Block header at the start of `src` through the quick parser and the portable one:
    -> ((bits, hlit, hdist, lens) or None when the quick parser declines, (error, bits, hlit, hdist, lens))."""
    L = lib()
    comp = len(src) if comp_bytes is None else comp_bytes
    lf, lp = ctypes.create_string_buffer(320), ctypes.create_string_buffer(320)
    out = (ctypes.c_uint32 * 7)()
    L.model_header_both(bytes(src), len(src), comp, lf, lp, out)
    fast = None if out[0] == 0 else (out[0], out[1], out[2], lf.raw[: out[1] + out[2]])
    port = (out[3], out[4], out[5], out[6], lp.raw[: out[5] + out[6]] if out[3] == 0 else b"")
    return fast, port
