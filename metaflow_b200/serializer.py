"""Scatter pickling: the protocol-4 pickle of an artifact as a list of segments instead of one bytes.

The reference serialises every artifact with `pickle.dumps(obj, protocol=4)`
(metaflow/datastore/task_datastore.py:351-355) -- for a 16 MiB array that is 10 ms of memcpy under
the GIL (ndarray.tobytes, then the pickler's growing output buffer) before a single byte is hashed.
Here the pickle STREAM is byte-for-byte the same (same SHA-1 key, same `size` in the task's
data.json), but the payload of every large C-contiguous ndarray is never materialised: the
pickler sees a 64 KiB placeholder in the array's place, and the emitted stream is cut around it.
The result is `Segments`: small `bytes` pieces plus zero-copy views of the arrays themselves,
which the codec copies straight into the page-locked staging arena (GIL released, in parallel).

Why the bytes are identical: protocol 4 writes any bytes payload >= 64 KiB outside the frames as
`BINBYTES <u32 length> <payload>` (CPython Modules/_pickle.c, _Pickler_write_bytes); nothing else in
the stream depends on the payload or on its length, so patching the four length bytes and splicing
the real buffer gives exactly what `pickle.dumps` would have produced.  Arrays below 64 KiB or of
4 GiB and more (different opcode / framing), non-contiguous, object-dtype or subclassed arrays take
the pickler's normal route.  `dumps_segments` verifies the splice points it finds and falls back to
`pickle.dumps` if anything about the stream is not as expected.
"""
import pickle
import struct

import numpy as np

_PLACEHOLDER_LEN = 1 << 16          # == pickle's FRAME_SIZE_TARGET: the smallest payload written out of frame
_PLACEHOLDER_HDR = b"B" + struct.pack("<I", _PLACEHOLDER_LEN)
_MAX_FAST = 1 << 32                 # BINBYTES8 from here on


class Segments(object):
    """An immutable byte string stored as pieces (bytes objects and uint8 ndarray views)."""

    __slots__ = ("parts", "nbytes")

    def __init__(self, parts):
        self.parts = parts
        self.nbytes = sum(len(p) for p in parts)

    def __len__(self):
        return self.nbytes

    def copy_into(self, dst, off):
        """dst: 1-D uint8 ndarray; writes the bytes at dst[off:off+len(self)]."""
        for p in self.parts:
            n = len(p)
            if n:
                dst[off:off + n] = p if isinstance(p, np.ndarray) else np.frombuffer(p, dtype=np.uint8)
                off += n

    @property
    def views(self):
        """The pieces as buffers, for writers that take a list of extents (LocalStorage._atomic_write)."""
        return self.parts

    def tobytes(self):
        return b"".join(p.tobytes() if isinstance(p, np.ndarray) else p for p in self.parts)

    __bytes__ = tobytes


class _Sink(object):
    """File object handed to the pickler: keeps what it is given, by reference."""

    __slots__ = ("parts", "pending", "ok")

    def __init__(self):
        self.parts = []
        self.pending = {}       # id(placeholder) -> (placeholder, uint8 view of the array)
        self.ok = True

    def write(self, b):
        ent = self.pending.pop(id(b), None) if self.pending else None
        if ent is None or ent[0] is not b:
            self.parts.append(b)
            return len(b)
        # the placeholder itself, handed over out of band: the BINBYTES header is the tail of what came before
        prev = self.parts[-1] if self.parts else b""
        if not (isinstance(prev, bytes) and prev.endswith(_PLACEHOLDER_HDR)):
            self.ok = False
            self.parts.append(b)
            return len(b)
        view = ent[1]
        self.parts[-1] = prev[:-4] + struct.pack("<I", len(view))
        self.parts.append(view)
        return len(b)


class _ScatterPickler(pickle.Pickler):
    def __init__(self, sink):
        super().__init__(sink, protocol=4)
        self._sink = sink

    def reducer_override(self, obj):
        if type(obj) is not np.ndarray:
            return NotImplemented
        nbytes = obj.nbytes
        if nbytes < _PLACEHOLDER_LEN or nbytes >= _MAX_FAST or not obj.flags.c_contiguous or obj.dtype.hasobject:
            return NotImplemented
        try:
            view = obj.reshape(-1).view(np.uint8)
            # numpy's own reduce on an empty array of the same dtype: same callable, same argument objects
            # (and object identities, which the pickler's memo sees) as the real array would give
            func, args, state = np.empty(0, dtype=obj.dtype).__reduce__()
        except Exception:
            return NotImplemented
        if len(view) != nbytes or len(state) != 5 or state[3] is not False:
            return NotImplemented
        placeholder = bytes(_PLACEHOLDER_LEN)
        self._sink.pending[id(placeholder)] = (placeholder, view)
        return func, args, (state[0], obj.shape, obj.dtype, False, placeholder)


def dumps_segments(obj):
    """-> Segments whose bytes equal pickle.dumps(obj, protocol=4).  Raises what pickle.dumps raises."""
    sink = _Sink()
    _ScatterPickler(sink).dump(obj)
    if not sink.ok or sink.pending:
        return Segments([pickle.dumps(obj, protocol=4)])
    return Segments(sink.parts)


# ----------------------------------------------------------------------------------------------------------
# Gather unpickling: the inverse.  pickle.loads of a 16 MiB array stream is one 16 MiB memcpy into a fresh
# bytes object (which numpy then adopts as the array's buffer) -- under the GIL, page-faulting 4096 new pages
# on the way: ~6 ms per artifact, serial.  For a BATCH of decoded blobs the payload copies are done first, by a
# thread pool with the GIL released (ctypes.memmove into not-yet-visible bytes objects), and the unpickler is
# then fed a rewritten stream in which every out-of-frame `BINBYTES <len> <payload>` is replaced by a
# persistent-id reference to the prepared object.  The objects that come out are the ones pickle.loads gives.
import ctypes
import io

_FRAME, _BINBYTES, _BINBYTES8, _PROTO = 0x95, 0x42, 0x8E, 0x80
_GATHER_MIN = 1 << 16

_py = ctypes.pythonapi
_new_bytes = _py.PyBytes_FromStringAndSize
_new_bytes.restype = ctypes.py_object
_new_bytes.argtypes = [ctypes.c_char_p, ctypes.c_ssize_t]
_bytes_addr = _py.PyBytes_AsString
_bytes_addr.restype = ctypes.c_void_p
_bytes_addr.argtypes = [ctypes.py_object]


_HUGE = 2 << 20
try:
    _libc = ctypes.CDLL(None, use_errno=True)
    _madvise = _libc.madvise
    _madvise.restype = ctypes.c_int
    _madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
except Exception:  # noqa: BLE001 -- no libc madvise: plain 4 KiB faults
    _madvise = None


def _want_huge_pages(addr, nbytes):
    """MADV_HUGEPAGE on the 2 MiB-aligned interior of a fresh, untouched buffer: a 16 MiB payload then takes 7
    page faults instead of 4096 when it is filled (transparent_hugepage=madvise is the common default)."""
    if _madvise is None or nbytes < 2 * _HUGE:
        return
    lo = (addr + _HUGE - 1) & ~(_HUGE - 1)
    hi = (addr + nbytes) & ~(_HUGE - 1)
    if hi > lo:
        _madvise(ctypes.c_void_p(lo), hi - lo, 14)          # 14 = MADV_HUGEPAGE; failure is harmless


def _walk(mv):
    """Top-level structure of a protocol-4 stream: [(is_payload, start, end)] covering mv[2:], or None when the
    stream is not PROTO 4 + frames + out-of-frame bytes payloads (then pickle.loads handles it)."""
    n = len(mv)
    if n < 4 or mv[0] != _PROTO or mv[1] != 4:
        return None
    out, pos, payloads = [], 2, 0
    while pos < n:
        op = mv[pos]
        if op == _FRAME:
            if pos + 9 > n:
                return None
            ln = int.from_bytes(mv[pos + 1:pos + 9], "little")
            end = pos + 9 + ln
            if end > n:
                return None
            out.append((False, pos, end))
            pos = end
        elif op == _BINBYTES or op == _BINBYTES8:
            h = 5 if op == _BINBYTES else 9
            if pos + h > n:
                return None
            ln = int.from_bytes(mv[pos + 1:pos + h], "little")
            end = pos + h + ln
            if end > n or ln < _GATHER_MIN:
                return None
            out.append((True, pos + h, end))
            payloads += 1
            pos = end
        else:
            return None
    return out if payloads else None


class _GatherUnpickler(pickle.Unpickler):
    def __init__(self, stream, objects):
        super().__init__(stream)
        self._objects = objects

    def persistent_load(self, pid):
        return self._objects[pid]


def loads_batch(blobs, pool=None):
    """[bytes-like protocol-4 pickle] -> [object], equal to [pickle.loads(b) for b in blobs].
    `pool`: a concurrent.futures executor for the payload copies (None: copy inline)."""
    plans, jobs = [], []
    for blob in blobs:
        mv = memoryview(blob)
        if mv.format != "B" or mv.ndim != 1:
            mv = mv.cast("B")
        segs = _walk(mv)
        if segs is None:
            plans.append(None)
            continue
        base = np.frombuffer(mv, dtype=np.uint8).ctypes.data
        objects, pieces = [], [bytes(mv[0:2])]
        for is_payload, a, b in segs:
            if is_payload:
                obj = _new_bytes(None, b - a)            # uninitialised; filled before anyone can see it
                dst = _bytes_addr(obj)
                _want_huge_pages(dst, b - a)
                jobs.append((dst, base + a, b - a))
                pieces.append(b"J" + struct.pack("<i", len(objects)) + b"Q")      # BININT idx, BINPERSID
                objects.append(obj)
            else:
                pieces.append(mv[a:b])
        plans.append((b"".join(pieces), objects))
    if jobs:
        def fill(job):
            ctypes.memmove(job[0], job[1], job[2])       # foreign call: the GIL is released
        if pool is not None and len(jobs) > 1:
            list(pool.map(fill, jobs))
        else:
            for job in jobs:
                fill(job)
    out = []
    for blob, plan in zip(blobs, plans):
        if plan is None:
            out.append(pickle.loads(blob))
        else:
            out.append(_GatherUnpickler(io.BytesIO(plan[0]), plan[1]).load())
    return out
