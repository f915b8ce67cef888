"""Round-robin sharding of a step's artifact set over the GPUs / ranks of one box.

North-star sharding rule: blob i of the batch belongs to shard i mod G; blobs are
independent, so there is no collective on the data path -- the only exchange is the host-side gather of the
small per-blob results (keys, sizes), in input order because callers rely on it
(metaflow/datastore/task_datastore.py:382).  Inside one process libmfcas does this over the devices of its
context; across processes (one rank per GPU, as bench.py runs under torchrun) these helpers do.
"""


def shard_indices(n_items, rank, world):
    """Indices of the items shard `rank` of `world` owns: rank, rank + world, ..."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_items, world))


def merge_shard_results(n_items, per_rank_results):
    """per_rank_results[r] = results of shard r in shard order -> one list in input order."""
    world = len(per_rank_results)
    out = [None] * n_items
    for r, res in enumerate(per_rank_results):
        idx = shard_indices(n_items, r, world)
        if len(idx) != len(res):
            raise ValueError("shard %d returned %d results for %d items" % (r, len(res), len(idx)))
        for i, v in zip(idx, res):
            out[i] = v
    return out


def pack_sharded(codec, blobs, rank=0, world=1, all_gather=None):
    """Pack this rank's shard with `codec`; with all_gather (e.g. torch.distributed.all_gather_object bound to
    a list) the keys of all ranks are merged back to input order."""
    mine = shard_indices(len(blobs), rank, world)
    keys, packed = codec.pack_batch([blobs[i] for i in mine])
    if all_gather is None or world == 1:
        return keys, packed
    gathered = all_gather(keys)
    return merge_shard_results(len(blobs), gathered), packed


# ------------------------------------------------------------------------------------------------------
# The multi-process batcher: ONE process per box owns the GPU context(s); task processes -- the children
# multicore_utils.parallel_map forks, or stock one-process-per-task runs (metaflow/runtime.py:1517-1577,
# 2287-2350) pointed at a running daemon -- hand their blobs over in anonymous shared memory (memfd; the
# DESCRIPTOR travels with the message, SCM_RIGHTS: no names in /dev/shm, nothing to leak when a client dies)
# and small JSON messages on a Unix socket only its owner can connect to.  Requests that arrive while a GPU batch is running are
# merged into the next one, so N concurrently saving tasks become one wide batch (a SHA-1 lane does
# ~150 MB/s; width is what makes the GPU path fast), sharded round-robin over the GPUs by libmfcas.
#
#   server:  CodecServer(codec).start()        or   python -m metaflow_b200.batcher serve [--socket PATH]
#   client:  METAFLOW_GPU_CAS_SERVER=<socket path>  ->  codec.default_codec() returns a RemoteCodec
import json
import mmap
import os
import queue
import socket
import struct
import tempfile
import threading

SERVER_ENV = "METAFLOW_GPU_CAS_SERVER"
_SHM_DIR = "/dev/shm" if os.path.isdir("/dev/shm") else None
_MAX_MSG = 256 << 20          # a message is lengths and keys, never payload


def _send_msg(sock, obj, fds=()):
    """One message: 8-byte length (descriptors ride on it), then JSON.  Nothing that crosses the socket is unpickled."""
    data = json.dumps(obj).encode("utf-8")
    hdr = struct.pack("<Q", len(data))
    if fds:
        sent = socket.send_fds(sock, [hdr], list(fds))
        if sent < len(hdr):
            sock.sendall(hdr[sent:])
    else:
        sock.sendall(hdr)
    sock.sendall(data)


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(n - len(buf), 1 << 20))
        if not chunk:
            raise EOFError("peer closed the batcher socket")
        buf += chunk
    return bytes(buf)


def _recv_msg(sock):
    """-> (message, [descriptors that came with it]); the caller owns the descriptors."""
    head, fds, _flags, _addr = socket.recv_fds(sock, 8, 4)
    if not head:
        for fd in fds:
            os.close(fd)
        raise EOFError("peer closed the batcher socket")
    try:
        if len(head) < 8:
            head += _recv_exact(sock, 8 - len(head))
        (n,) = struct.unpack("<Q", head)
        if n > _MAX_MSG:
            raise ValueError("batcher message of %d bytes" % n)
        return json.loads(_recv_exact(sock, n).decode("utf-8")), list(fds)
    except Exception:
        for fd in fds:
            os.close(fd)
        raise


def _layout(lens):
    """Byte offsets of blobs laid out back to back at 64-byte aligned positions -> (offsets, total)."""
    offs, cur = [], 0
    for ln in lens:
        offs.append(cur)
        cur += (ln + 63) // 64 * 64
    return offs, cur


class _Segment(object):
    """Anonymous shared memory holding a list of byte strings back to back; its descriptor is what crosses the socket."""

    def __init__(self, fd, size):
        self.fd, self.size = fd, size
        try:
            self.map = mmap.mmap(fd, max(size, 1))
        except Exception:
            os.close(fd)
            self.fd = -1
            raise
        self.view = memoryview(self.map)

    @classmethod
    def create(cls, size):
        if hasattr(os, "memfd_create"):
            fd = os.memfd_create("mfcas_seg", getattr(os, "MFD_CLOEXEC", 0))
        else:                                   # no memfd: an unlinked file in shared memory is the same thing
            fd, path = tempfile.mkstemp(prefix="mfcas_seg_", dir=_SHM_DIR)
            os.unlink(path)
        try:
            os.ftruncate(fd, max(size, 1))
        except Exception:
            os.close(fd)
            raise
        return cls(fd, size)

    def close_fd(self):
        if self.fd >= 0:
            os.close(self.fd)
            self.fd = -1

    def close(self):
        self.close_fd()
        try:
            self.view.release()
            self.map.close()
        except (BufferError, ValueError):
            pass                       # views handed to the caller are still alive: the mapping goes with them


class _Request(object):
    __slots__ = ("op", "items", "done", "result", "segment", "error")

    def __init__(self, op, items):
        self.op, self.items = op, items
        self.done = threading.Event()
        self.result = self.segment = self.error = None


class CodecServer(object):
    """Owns `codec` (a GpuCodec) and serves pack / hash / unpack requests of other processes, merging
    concurrent requests of the same kind into one codec call."""

    def __init__(self, codec, socket_path=None, linger=0.002, max_batch_bytes=4 << 30):
        self.codec = codec
        self.linger = linger
        self.max_batch_bytes = max_batch_bytes
        self._own_dir = None
        if socket_path is None:
            self._own_dir = tempfile.mkdtemp(prefix="mfcas_srv_")
            socket_path = os.path.join(self._own_dir, "sock")
        self.socket_path = socket_path
        self._sock = None
        self._q = queue.Queue()
        self._stop = threading.Event()
        self._threads = []
        self.batches = {"pack": 0, "hash": 0, "unpack": 0}       # codec calls issued (tests read this)
        self.requests = 0

    # -- lifecycle
    def start(self):
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        old_umask = os.umask(0o177)              # the socket is the access control: owner only
        try:
            s.bind(self.socket_path)
        finally:
            os.umask(old_umask)
        os.chmod(self.socket_path, 0o600)
        s.listen(256)
        s.settimeout(0.2)
        self._sock = s
        for target in (self._accept_loop, self._batch_loop):
            t = threading.Thread(target=target, daemon=True, name="mfcas-" + target.__name__)
            t.start()
            self._threads.append(t)
        return self

    def stop(self):
        self._stop.set()
        self._q.put(None)
        for t in self._threads:
            t.join(timeout=5)
        if self._sock is not None:
            self._sock.close()
        try:
            os.unlink(self.socket_path)
        except OSError:
            pass
        if self._own_dir is not None:
            try:
                os.rmdir(self._own_dir)
            except OSError:
                pass

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()

    # -- connections: one thread per client, requests answered in order
    def _accept_loop(self):
        while not self._stop.is_set():
            try:
                conn, _ = self._sock.accept()
            except socket.timeout:
                continue
            except OSError:
                return
            threading.Thread(target=self._serve, args=(conn,), daemon=True, name="mfcas-conn").start()

    def _serve(self, conn):
        with conn:
            while not self._stop.is_set():
                try:
                    msg, fds = _recv_msg(conn)
                except (EOFError, OSError, ValueError):
                    return
                seg = out_seg = None
                try:
                    op = msg["op"]
                    if op == "ping":
                        _send_msg(conn, {"ok": True, "pid": os.getpid()})
                        continue
                    items = []
                    if msg.get("size") is not None and fds:
                        seg = _Segment(fds.pop(0), int(msg["size"]))
                        lens = [int(x) for x in msg["lens"]]
                        offs, total = _layout(lens)
                        if total > seg.size:
                            raise ValueError("segment of %d bytes for %d bytes of blobs" % (seg.size, total))
                        items = [seg.view[o:o + ln] for o, ln in zip(offs, lens)]
                    # unpack requests may name files instead: the server reads them straight into pinned memory
                    for k, path in sorted((int(k), v) for k, v in msg.get("files", {}).items()):
                        items.insert(k, self._file_source(path))
                    req = _Request(op, items)
                    self._q.put(req)
                    req.done.wait()
                    out_seg = req.segment
                    if req.error is not None:
                        _send_msg(conn, {"ok": False, "error": req.error})
                    else:
                        _send_msg(conn, req.result, [out_seg.fd] if out_seg is not None else ())
                except (OSError, KeyError, ValueError, TypeError) as e:
                    try:
                        _send_msg(conn, {"ok": False, "error": "%s: %s" % (type(e).__name__, e)})
                    except OSError:
                        return
                finally:
                    for fd in fds:                   # descriptors the message did not account for
                        os.close(fd)
                    if out_seg is not None:          # the client has its own descriptor now (or is gone): ours can go
                        out_seg.close()
                    if seg is not None:
                        items = None
                        seg.close()

    def _file_source(self, path):
        if getattr(self.codec, "accepts_file_sources", False):
            from .codec import FileSource

            return FileSource(path, os.path.getsize(path))
        with open(path, "rb") as f:
            return f.read()

    # -- the batcher proper
    def _batch_loop(self):
        while True:
            first = self._q.get()
            if first is None:
                return
            batch, size = [first], sum(len(b) for b in first.items)
            deadline = self.linger
            held = []
            while size < self.max_batch_bytes:
                try:
                    nxt = self._q.get(timeout=deadline)
                except queue.Empty:
                    break
                if nxt is None:
                    self._q.put(None)
                    break
                if nxt.op != first.op:
                    held.append(nxt)
                    continue
                batch.append(nxt)
                size += sum(len(b) for b in nxt.items)
                deadline = 0.0005
            for h in held:
                self._q.put(h)
            self._run(first.op, batch)

    def _run(self, op, batch):
        self.requests += len(batch)
        flat = [b for r in batch for b in r.items]
        try:
            if op == "hash":
                keys = self.codec.hash_batch(flat) if flat else []
                self.batches["hash"] += 1
                pos = 0
                for r in batch:
                    r.result = {"ok": True, "keys": keys[pos:pos + len(r.items)]}
                    pos += len(r.items)
            elif op == "pack":
                keys, packed = self.codec.pack_batch(flat) if flat else ([], [])
                self.batches["pack"] += 1
                pos = 0
                for r in batch:
                    mine = packed[pos:pos + len(r.items)]
                    meta, r.segment = self._export([getattr(p, "views", [p]) for p in mine])
                    r.result = dict(meta, ok=True, keys=keys[pos:pos + len(r.items)])
                    pos += len(r.items)
            elif op == "unpack":
                out = self.codec.unpack_batch(flat, as_bytes=False) if flat else []
                self.batches["unpack"] += 1
                pos = 0
                for r in batch:
                    mine = out[pos:pos + len(r.items)]
                    errors = {str(i): [getattr(o, "code", -1), str(o)] for i, o in enumerate(mine) if isinstance(o, Exception)}
                    meta, r.segment = self._export([[o] if not isinstance(o, Exception) else [] for o in mine])
                    r.result = dict(meta, ok=True, errors=errors)
                    pos += len(r.items)
            else:
                raise ValueError("unknown op %r" % (op,))
        except Exception as e:  # noqa: BLE001 -- a failed batch is reported to every client that was part of it
            for r in batch:
                if r.segment is not None:
                    r.segment.close()
                r.result, r.segment, r.error = None, None, "%s: %s" % (type(e).__name__, e)
        for r in batch:
            r.done.set()

    @staticmethod
    def _export(pieces_per_blob):
        """[[bytes-like piece, ...] per blob] -> (message fields, result segment): every blob's pieces joined, back to back.
        The segment's descriptor goes out with the reply; the caller closes the segment afterwards."""
        lens = [sum(len(p) for p in pieces) for pieces in pieces_per_blob]
        offs, total = _layout(lens)
        seg = _Segment.create(total)
        try:
            for o, pieces in zip(offs, pieces_per_blob):
                cur = o
                for p in pieces:
                    seg.view[cur:cur + len(p)] = p
                    cur += len(p)
        except Exception:
            seg.close()
            raise
        return {"size": total, "lens": lens}, seg


class RemoteCodec(object):
    """The GpuCodec interface (hash_batch / pack_batch / unpack_batch) served by a CodecServer in another
    process.  Results are views over a shared-memory mapping that lives as long as they do."""

    accepts_file_sources = True

    def __init__(self, socket_path=None):
        from .exception import GpuCasUnavailable

        self.socket_path = socket_path or os.environ.get(SERVER_ENV)
        if not self.socket_path:
            raise GpuCasUnavailable("no batcher address: set %s" % SERVER_ENV)
        try:
            self._sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            self._sock.connect(self.socket_path)
        except OSError as e:
            raise GpuCasUnavailable("cannot reach the GPU batcher at %s: %s" % (self.socket_path, e))
        self._lock = threading.Lock()

    def close(self):
        try:
            self._sock.close()
        except OSError:
            pass

    def ping(self):
        with self._lock:
            _send_msg(self._sock, {"op": "ping"})
            return _recv_msg(self._sock)[0]

    def _call(self, op, blobs):
        from .codec import FileSource
        from .exception import DataException
        from .serializer import Segments

        files = {str(i): b.path for i, b in enumerate(blobs) if isinstance(b, FileSource)}
        mem = [b for b in blobs if not isinstance(b, FileSource)]
        lens = [len(b) for b in mem]
        offs, total = _layout(lens)
        seg = None
        msg = {"op": op, "files": files}
        fds = []
        try:
            if mem:
                seg = _Segment.create(total)
                import numpy as np

                arr = np.frombuffer(seg.view, dtype=np.uint8)
                for o, ln, b in zip(offs, lens, mem):
                    if not ln:
                        continue
                    if isinstance(b, Segments):
                        b.copy_into(arr, o)
                    else:
                        seg.view[o:o + ln] = b
                del arr
                msg.update(size=total, lens=lens)
            with self._lock:
                _send_msg(self._sock, msg, [seg.fd] if seg is not None else ())
                rep, fds = _recv_msg(self._sock)
        finally:
            if seg is not None:
                seg.close()
        if not rep.get("ok"):
            for fd in fds:
                os.close(fd)
            raise DataException("GPU batcher call failed: %s" % rep.get("error"))
        return rep, fds

    @staticmethod
    def _import(rep, fds):
        """The reply's segment -> one view per blob; the mapping lives as long as the views do."""
        if not fds:
            raise EOFError("the batcher's reply came without its segment")
        for fd in fds[1:]:
            os.close(fd)
        out_seg = _Segment(fds[0], int(rep["size"]))
        out_seg.close_fd()                                 # the mapping keeps the pages
        lens = [int(x) for x in rep["lens"]]
        offs, total = _layout(lens)
        if total > out_seg.size:
            raise ValueError("reply segment of %d bytes for %d bytes of blobs" % (out_seg.size, total))
        return [out_seg.view[o:o + ln] for o, ln in zip(offs, lens)]

    def hash_batch(self, blobs):
        if not blobs:
            return []
        rep, fds = self._call("hash", list(blobs))
        for fd in fds:
            os.close(fd)
        return rep["keys"]

    def pack_batch(self, blobs):
        from .codec import Packed

        if not blobs:
            return [], []
        rep, fds = self._call("pack", list(blobs))
        return rep["keys"], [Packed([v]) for v in self._import(rep, fds)]

    def unpack_batch(self, packed_blobs, as_bytes=True):
        from .codec import UnpackError

        if not packed_blobs:
            return []
        rep, fds = self._call("unpack", list(packed_blobs))
        views = self._import(rep, fds)
        errors = {int(k): v for k, v in rep["errors"].items()}
        out = []
        for i, v in enumerate(views):
            if i in errors:
                out.append(UnpackError(errors[i][0]))
            else:
                out.append(bytes(v) if as_bytes else v)
        return out


def serve_forever(socket_path, devices=None):
    from .codec import GpuCodec

    server = CodecServer(GpuCodec(devices=devices), socket_path=socket_path).start()
    print("mfcas batcher serving on %s (export %s=%s)" % (server.socket_path, SERVER_ENV, server.socket_path), flush=True)
    try:
        threading.Event().wait()
    except KeyboardInterrupt:
        pass
    finally:
        server.stop()


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="GPU artifact batcher daemon")
    ap.add_argument("cmd", choices=["serve"])
    ap.add_argument("--socket", default=os.path.join(tempfile.gettempdir(), "mfcas_batcher_%d.sock" % os.getuid()))
    ap.add_argument("--devices", default="")
    a = ap.parse_args()
    serve_forever(a.socket, [int(x) for x in a.devices.split(",") if x] or None)
