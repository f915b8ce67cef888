"""GpuCodec: the batched blob codec behind ContentAddressedStore (pack = SHA-1 + gzip, unpack = gunzip).

It replaces, for a whole batch at a time, the three per-blob calls of the reference
(metaflow/datastore/content_addressed_store.py:98 sha1, :114/:211-216 _pack_v1, :192/:218-220
_unpack_v1) with mfcas_pack_batch / mfcas_hash_batch / mfcas_unpack_batch on the GPUs.
Blobs are packed into page-locked arenas so every wave moves with one cudaMemcpyAsync each way.
"""
import ctypes
import os
import struct
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import capi
from .exception import DataException, GpuCasUnavailable
from .serializer import Segments
from . import metaflow_config

_ALIGN = 256


def _align(n):
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class _PinnedArena(object):
    """A grow-only page-locked byte buffer exposed as a numpy uint8 array."""

    def __init__(self, lib):
        self._lib = lib
        self._ptr = None
        self.cap = 0
        self.arr = None

    def ensure(self, nbytes):
        if nbytes <= self.cap:
            return
        self.release()
        want = int(nbytes + nbytes // 8 + 4096)
        p = self._lib.mfcas_alloc_pinned(want)
        if not p:
            raise MemoryError("mfcas_alloc_pinned(%d) failed" % want)
        self._ptr = p
        self.cap = want
        self.arr = np.ctypeslib.as_array((ctypes.c_uint8 * want).from_address(p))

    @property
    def base(self):
        return self._ptr

    def release(self):
        if self._ptr:
            self.arr = None
            self._lib.mfcas_free_pinned(self._ptr)
            self._ptr = None
            self.cap = 0

    def trim(self, keep):
        """Give page-locked memory back after an oversized window (the arena is grow-only otherwise)."""
        if self.cap > keep:
            self.release()


_MF_HDR_FIXED, _MF_TRAILER = 32, 13


def packed_size_hint(path, size):
    """Decoded size of a packed CAS file WITHOUT reading its payload: exact for files this library wrote (the
    members' index headers are walked with a few small reads), the ISIZE trailer (mod 4 GiB, last member only)
    for foreign gzip, the file size for anything else.  Used to bound load windows by UNPACKED bytes."""
    try:
        with open(path, "rb", buffering=0) as f:
            pos, total, members = 0, 0, 0
            while pos < size:
                f.seek(pos)
                h = f.read(_MF_HDR_FIXED)
                if len(h) < _MF_HDR_FIXED or h[:4] != b"\x1f\x8b\x08\x04" or h[12:14] != b"MF":
                    break
                raw_len, nc = struct.unpack_from("<II", h, 20)
                idx = f.read(8 * nc)
                if len(idx) < 8 * nc:
                    break
                payload = sum(struct.unpack_from("<I", idx, 8 * k)[0] for k in range(nc))
                pos += _MF_HDR_FIXED + 72 * nc + payload + _MF_TRAILER
                total += raw_len
                members += 1
            if members and pos >= size:
                return total
            if size >= 18:
                f.seek(size - 4)
                (isize,) = struct.unpack("<I", f.read(4))
                return max(total + isize, size)
    except OSError:
        pass
    return size


class FileSource(object):
    """A packed blob that still lives in a file: staged with readinto() straight into the pinned arena."""

    __slots__ = ("path", "size")

    def __init__(self, path, size):
        self.path = path
        self.size = int(size)

    def __len__(self):
        return self.size

    def read_into(self, dst, off):
        view = memoryview(dst)[off:off + self.size]
        got = 0
        with open(self.path, "rb", buffering=0) as f:
            while got < self.size:
                k = f.readinto(view[got:])
                if not k:
                    raise DataException("'%s' is shorter than its size on open (%d of %d bytes)" % (self.path, got, self.size))
                got += k

    def head(self, n=65536):
        with open(self.path, "rb", buffering=0) as f:
            return f.read(min(n, self.size))


def _as_u8(blob):
    if isinstance(blob, np.ndarray) and blob.dtype == np.uint8 and blob.ndim == 1:
        return blob
    return np.frombuffer(blob, dtype=np.uint8)


class Packed(object):
    """The RFC 1952 bytes of one packed blob as one or more views over the page-locked output arena
    (several when the blob was packed in stripes: the file is their concatenation -- writers writev them).
    Valid until the next pack/unpack call of the codec that produced it."""

    __slots__ = ("views",)

    def __init__(self, views):
        self.views = views

    def __len__(self):
        return sum(len(v) for v in self.views)

    def __bytes__(self):
        return bytes(self.views[0]) if len(self.views) == 1 else b"".join(bytes(v) for v in self.views)

    def tobytes(self):
        return bytes(self)


class UnpackError(Exception):
    """Per-blob decode failure; str() mirrors the message the reference's gzip reader gives."""

    def __init__(self, code):
        self.code = code
        super().__init__(capi.STATUS_TEXT.get(code, "error %d" % code))


def _locked(method):
    def guarded(self, *a, **kw):
        with self._lock:
            return method(self, *a, **kw)

    guarded.__name__, guarded.__doc__ = method.__name__, method.__doc__
    return guarded


class GpuCodec(object):
    def __init__(self, devices=None, wave_bytes=None, lz=None):
        lib = capi.load()
        if lib.mfcas_device_count() <= 0:
            raise GpuCasUnavailable("no CUDA device visible: the artifact hot path has no CPU fallback")
        if devices is None and metaflow_config.GPU_CAS_DEVICES:
            devices = [int(x) for x in str(metaflow_config.GPU_CAS_DEVICES).split(",") if x != ""]
        if wave_bytes is None:
            wave_bytes = int(metaflow_config.GPU_CAS_WAVE_BYTES)
        if lz is None:
            lz = bool(int(metaflow_config.GPU_CAS_LZ))
        flags = 0 if lz else capi.FLAG_NO_LZ
        if devices:
            arr = (ctypes.c_int * len(devices))(*devices)
            ctx = lib.mfcas_create(arr, len(devices), wave_bytes, flags)
        else:
            ctx = lib.mfcas_create(None, 0, wave_bytes, flags)
        if not ctx:
            raise GpuCasUnavailable("mfcas_create failed (no usable CUDA device or out of memory)")
        self._lib = lib
        self._ctx = ctypes.c_void_p(ctx)
        self._in = _PinnedArena(lib)
        self._out = _PinnedArena(lib)
        # The page-locked arenas are shared by every call: the views a call returns (Packed, decoded memoryviews)
        # point into them and die with the next call.  Callers that keep views across statements -- the
        # content-addressed store writes files from them, the task datastore unpickles from them -- hold
        # `session()` for as long as they use the views; every codec method takes the same (re-entrant) lock.
        self._lock = threading.RLock()

    def session(self):
        """Context manager: exclusive use of the codec and its arenas (re-entrant per thread)."""
        return self._lock

    def trim(self, keep=None):
        """Release page-locked arenas larger than `keep` bytes (default METAFLOW_GPU_CAS_ARENA_KEEP)."""
        keep = int(metaflow_config.GPU_CAS_ARENA_KEEP) if keep is None else keep
        with self._lock:
            self._in.trim(keep)
            self._out.trim(keep)

    def close(self):
        if getattr(self, "_ctx", None):
            self._in.release()
            self._out.release()
            self._lib.mfcas_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != capi.OK:
            raise DataException("GPU codec call failed (%s): %s" % (
                capi.STATUS_TEXT.get(rc, rc), self._lib.mfcas_last_error(self._ctx).decode("utf-8", "replace")))

    @staticmethod
    def _slots(sizes):
        """sizes (uint64 array) -> (offsets, total): 256-byte aligned slots with a 256-byte gap, vectorised."""
        step = (sizes + np.uint64(_ALIGN - 1)) // np.uint64(_ALIGN) * np.uint64(_ALIGN) + np.uint64(_ALIGN)
        ends = np.cumsum(step, dtype=np.uint64)
        return (ends - step).astype(np.uint64), int(ends[-1]) if len(ends) else 0

    def _stage_inputs(self, blobs):
        n = len(blobs)
        lens = np.fromiter((len(b) for b in blobs), dtype=np.uint64, count=n)
        offs, total = self._slots(lens)
        self._in.ensure(total + _ALIGN)
        arr = self._in.arr
        lens_l, offs_l = lens.tolist(), offs.tolist()

        def stage(lo, hi):
            for i in range(lo, hi):
                ln = lens_l[i]
                if ln:
                    o = offs_l[i]
                    b = blobs[i]
                    if isinstance(b, Segments):
                        b.copy_into(arr, o)                # pickle stream pieces + the arrays' own buffers
                    elif isinstance(b, FileSource):
                        b.read_into(arr, o)                # file -> page-locked memory, no bytes object in between
                    else:
                        arr[o:o + ln] = _as_u8(b)          # numpy releases the GIL for large copies

        if n >= 4 and total >= (64 << 20):
            # a few contiguous ranges of about equal bytes per worker (not one task per blob: a wide foreach of
            # small artifacts is 65536 blobs)
            workers = 8
            target = total // (workers * 4) + 1
            cuts, acc = [0], 0
            for i in range(n):
                acc += lens_l[i]
                if acc >= target:
                    cuts.append(i + 1)
                    acc = 0
            if cuts[-1] != n:
                cuts.append(n)
            with ThreadPoolExecutor(max_workers=workers) as pool:
                list(pool.map(lambda k: stage(cuts[k], cuts[k + 1]), range(len(cuts) - 1)))
        else:
            stage(0, n)
        src = (offs + np.uint64(self._in.base)).astype(np.uint64)
        return lens, src

    @staticmethod
    def _pp(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p))

    @staticmethod
    def _p64(a):
        return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))

    # ------------------------------------------------------------------ save side
    @_locked
    def hash_batch(self, blobs):
        """-> list of 40-hex SHA-1 keys (content_addressed_store.py:98)."""
        n = len(blobs)
        if n == 0:
            return []
        lens, src = self._stage_inputs(blobs)
        sha = np.empty(20 * n, dtype=np.uint8)
        st = np.zeros(n, dtype=np.int32)
        self._check(self._lib.mfcas_hash_batch(self._ctx, n, self._pp(src), self._p64(lens), sha.ctypes.data,
                                               st.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        raw = sha.tobytes()
        return [raw[20 * i:20 * i + 20].hex() for i in range(n)]

    @_locked
    def pack_batch(self, blobs):
        """-> (keys, packed): packed[i] is a `Packed` (views over page-locked memory holding the RFC 1952
        bytes of blob i, one view per 2 MiB stripe).  Valid until the next pack/unpack call.
        Uses mfcas_pack_batch_extents: stripes of all blobs are uploaded, hashed, encoded and copied back
        wave by wave, so nothing trails the last blob."""
        n = len(blobs)
        if n == 0:
            return [], []
        lib = self._lib
        lens, src = self._stage_inputs(blobs)
        geo = {}                                   # a foreach's artifacts have few distinct lengths
        for x in set(lens.tolist()):
            geo[x] = (lib.mfcas_pack_extent_count(x), lib.mfcas_pack_extent_bound(x))
        lens_l = lens.tolist()
        counts = np.fromiter((geo[x][0] for x in lens_l), dtype=np.uint64, count=n)
        caps = np.fromiter((geo[x][1] for x in lens_l), dtype=np.uint64, count=n)
        ooffs, total = self._slots(caps)
        self._out.ensure(total + _ALIGN)
        dst = (ooffs + np.uint64(self._out.base)).astype(np.uint64)
        n_ext = int(counts.sum())
        ext_off = np.zeros(n_ext, dtype=np.uint64)
        ext_len = np.zeros(n_ext, dtype=np.uint64)
        sha = np.empty(20 * n, dtype=np.uint8)
        st = np.zeros(n, dtype=np.int32)
        self._check(lib.mfcas_pack_batch_extents(self._ctx, n, self._pp(src), self._p64(lens), self._pp(dst), self._p64(caps),
                                                 self._p64(ext_off), self._p64(ext_len), sha.ctypes.data,
                                                 st.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        if st.any():
            bad = int(np.flatnonzero(st)[0])
            raise DataException("GPU pack failed for blob %d: %s" % (bad, capi.STATUS_TEXT.get(int(st[bad]), st[bad])))
        raw = sha.tobytes()
        keys = [raw[20 * i:20 * i + 20].hex() for i in range(n)]
        mv = memoryview(self._out.arr)
        starts = (np.repeat(ooffs, counts.astype(np.int64)) + ext_off).tolist()       # absolute start of every extent
        ends = (np.repeat(ooffs, counts.astype(np.int64)) + ext_off + ext_len).tolist()
        packed, k = [], 0
        for c in counts.tolist():
            packed.append(Packed([mv[starts[j]:ends[j]] for j in range(k, k + c)]))
            k += c
        return keys, packed

    @_locked
    def pack_batch_contiguous(self, blobs):
        """Same result through mfcas_pack_batch: one contiguous view per blob (blob-granular waves)."""
        n = len(blobs)
        if n == 0:
            return [], []
        lens, src = self._stage_inputs(blobs)
        bound = {x: self._lib.mfcas_pack_bound(x) for x in set(lens.tolist())}
        caps = np.fromiter((bound[x] for x in lens.tolist()), dtype=np.uint64, count=n)
        ooffs, total = self._slots(caps)
        self._out.ensure(total + _ALIGN)
        dst = (ooffs + np.uint64(self._out.base)).astype(np.uint64)
        dlen = np.zeros(n, dtype=np.uint64)
        sha = np.empty(20 * n, dtype=np.uint8)
        st = np.zeros(n, dtype=np.int32)
        self._check(self._lib.mfcas_pack_batch(self._ctx, n, self._pp(src), self._p64(lens), self._pp(dst), self._p64(caps),
                                               self._p64(dlen), sha.ctypes.data, st.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        if st.any():
            bad = int(np.flatnonzero(st)[0])
            raise DataException("GPU pack failed for blob %d: %s" % (bad, capi.STATUS_TEXT.get(int(st[bad]), st[bad])))
        raw = sha.tobytes()
        keys = [raw[20 * i:20 * i + 20].hex() for i in range(n)]
        mv = memoryview(self._out.arr)
        return keys, [Packed([mv[int(ooffs[i]):int(ooffs[i]) + int(dlen[i])]]) for i in range(n)]

    # ------------------------------------------------------------------ load side
    def unpacked_size(self, packed):
        a = _as_u8(packed)
        size = ctypes.c_uint64(0)
        exact = ctypes.c_int(0)
        rc = self._lib.mfcas_unpacked_size(a.ctypes.data, a.size, ctypes.byref(size), ctypes.byref(exact))
        if rc != capi.OK:
            raise UnpackError(rc)
        return size.value, bool(exact.value)

    def _unpacked_size_at(self, ptr, nbytes):
        size = ctypes.c_uint64(0)
        exact = ctypes.c_int(0)
        rc = self._lib.mfcas_unpacked_size(ctypes.c_void_p(ptr), nbytes, ctypes.byref(size), ctypes.byref(exact))
        if rc != capi.OK:
            raise UnpackError(rc)
        return size.value

    accepts_file_sources = True

    @_locked
    def unpack_batch(self, packed_blobs, as_bytes=True):
        """-> list with, per blob, the decoded bytes (or a memoryview over pinned memory when
        as_bytes=False, valid until the next call) or an UnpackError instance.
        Items are bytes-like or `FileSource`s (read straight into the page-locked staging arena)."""
        n = len(packed_blobs)
        if n == 0:
            return []
        results = [None] * n
        try:
            all_lens, all_src = self._stage_inputs(packed_blobs)       # staged once, also for retries
        except OSError as e:
            raise DataException("could not read a packed blob: %s" % e)
        sizes = np.zeros(n, dtype=np.uint64)
        todo = []
        for i in range(n):
            try:
                sizes[i] = self._unpacked_size_at(int(all_src[i]), int(all_lens[i]))
                todo.append(i)
            except UnpackError as e:
                results[i] = e
        attempt = 0
        while todo:
            idx = np.array(todo, dtype=np.int64)
            lens, src = np.ascontiguousarray(all_lens[idx]), np.ascontiguousarray(all_src[idx])
            m = len(todo)
            caps = np.ascontiguousarray(sizes[idx])
            ooffs, total = self._slots(caps)
            self._out.ensure(total + _ALIGN)
            dst = (ooffs + np.uint64(self._out.base)).astype(np.uint64)
            dlen = np.zeros(m, dtype=np.uint64)
            st = np.zeros(m, dtype=np.int32)
            self._check(self._lib.mfcas_unpack_batch(self._ctx, m, self._pp(src), self._p64(lens), self._pp(dst), self._p64(caps),
                                                     self._p64(dlen), st.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
            mv = memoryview(self._out.arr)
            retry = []
            st_l, oo_l, dl_l = st.tolist(), ooffs.tolist(), dlen.tolist()
            for k, i in enumerate(todo):
                code = st_l[k]
                if code == capi.OK:
                    view = mv[oo_l[k]:oo_l[k] + dl_l[k]]
                    results[i] = bytes(view) if (as_bytes or retry) else view
                elif code == capi.E_DST_TOO_SMALL and attempt < 8:
                    # foreign multi-member / >4 GiB streams: the ISIZE hint was short -- grow and retry
                    sizes[i] = max(dl_l[k], 2 * int(sizes[i]) + 65536)
                    retry.append(i)
                else:
                    results[i] = UnpackError(code)
            if retry and not as_bytes:
                # views of this round die when the arena is reused: materialise them
                for i in todo:
                    if isinstance(results[i], memoryview):
                        results[i] = bytes(results[i])
            todo = retry
            attempt += 1
        return results

    def counters(self):
        a, b, c = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        self._lib.mfcas_counters(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"kernel_launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}


_default = None
_forked_from_gpu_parent = False


def default_codec():
    """Process-wide codec; raises GpuCasUnavailable when the GPU path cannot run.  With
    METAFLOW_GPU_CAS_SERVER set (a batcher daemon, or a parent that forked us through
    multicore_utils.parallel_map) it is a RemoteCodec: this process never creates a CUDA context."""
    global _default
    if _default is None:
        from . import batcher

        if os.environ.get(batcher.SERVER_ENV):
            _default = batcher.RemoteCodec()
        elif _forked_from_gpu_parent:
            raise GpuCasUnavailable("forked from a process that owns the GPU context and no batcher is reachable")
        else:
            _default = GpuCodec()
    return _default
