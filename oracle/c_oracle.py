"""ctypes loader for oracle/mfcas_oracle.c (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmfcas_oracle.so")


def build():
    src = os.path.join(_HERE, "mfcas_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        u8p = ctypes.c_char_p
        L.mfo_sha1.argtypes = [u8p, ctypes.c_size_t, ctypes.c_void_p]
        L.mfo_sha1.restype = None
        L.mfo_crc32.argtypes = [ctypes.c_uint32, u8p, ctypes.c_size_t]
        L.mfo_crc32.restype = ctypes.c_uint32
        L.mfo_inflate.argtypes = [u8p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                  ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
        L.mfo_inflate.restype = ctypes.c_int
        L.mfo_gunzip.argtypes = [u8p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                 ctypes.POINTER(ctypes.c_size_t)]
        L.mfo_gunzip.restype = ctypes.c_int
        L.mfo_gzip_stored_bound.argtypes = [ctypes.c_size_t]
        L.mfo_gzip_stored_bound.restype = ctypes.c_size_t
        L.mfo_gzip_stored.argtypes = [u8p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_void_p]
        L.mfo_gzip_stored.restype = ctypes.c_size_t
        _lib = L
    return _lib


def sha1(data):
    out = ctypes.create_string_buffer(20)
    lib().mfo_sha1(bytes(data), len(data), out)
    return out.raw


def crc32(data, crc=0):
    return lib().mfo_crc32(crc, bytes(data), len(data))


def gunzip(data, cap):
    """Returns (errcode, bytes)."""
    out = ctypes.create_string_buffer(max(cap, 1))
    n = ctypes.c_size_t(0)
    e = lib().mfo_gunzip(bytes(data), len(data), out, cap, ctypes.byref(n))
    return e, out.raw[: n.value] if e == 0 else b""


def inflate(data, cap):
    out = ctypes.create_string_buffer(max(cap, 1))
    used = ctypes.c_size_t(0)
    n = ctypes.c_size_t(0)
    e = lib().mfo_inflate(bytes(data), len(data), out, cap, ctypes.byref(used), ctypes.byref(n))
    return e, used.value, out.raw[: n.value] if e == 0 else b""


def gzip_stored(data, mtime=0):
    cap = lib().mfo_gzip_stored_bound(len(data))
    out = ctypes.create_string_buffer(cap)
    n = lib().mfo_gzip_stored(bytes(data), len(data), mtime, out)
    return out.raw[:n]
