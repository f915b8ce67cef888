"""Content-addressed store with a batched GPU codec.

Same surface and on-disk format as metaflow/datastore/content_addressed_store.py:11-228
(`ContentAddressedStore(prefix, storage_impl)`, `.TYPE`, `.set_blob_cache`, `save_blobs` :41-120,
`load_blobs` :122-204, `_pack_v1` :211-216, `_unpack_v1` :218-220, `BlobCache` :223-228):
key = sha1(blob).hexdigest(), path = <prefix>/<key[:2]>/<key>, sidecar meta
{"cas_raw": raw, "cas_version": 1}, packed blobs are RFC 1952 files that stock Metaflow reads.

What changes is HOW: the reference hashes and gzips one blob at a time on one CPU core; here the
iterator is drained in windows, and each window goes through GpuCodec.pack_batch /
unpack_batch (SHA-1, DEFLATE, CRC-32, INFLATE on the GPUs, sharded round-robin across devices).
"""
import os
from collections import OrderedDict, namedtuple
from contextlib import nullcontext
from io import BytesIO

from ..exception import DataException
from .. import metaflow_config
from ..codec import packed_size_hint


class ContentAddressedStore(object):
    save_blobs_result = namedtuple("save_blobs_result", "uri key")

    def __init__(self, prefix, storage_impl, codec=None):
        self._prefix = prefix
        self._storage_impl = storage_impl
        self.TYPE = self._storage_impl.TYPE
        self._blob_cache = None
        self._codec = codec
        # bytes of raw blobs gathered before a GPU batch is issued
        self._window_bytes = int(metaflow_config.GPU_CAS_WINDOW_BYTES)
        self._window_unpacked = int(metaflow_config.GPU_CAS_WINDOW_UNPACKED_BYTES)
        self._window_blobs = int(metaflow_config.GPU_CAS_WINDOW_BLOBS)
        # content -> key of the tiny blobs this store has already keyed (the ~14 control artifacts every task of a
        # foreach saves: `True`, `None`, the step name ...): a repeat is answered without touching the GPU
        self._small_keys = OrderedDict()

    # -- plumbing ----------------------------------------------------------------------
    @property
    def codec(self):
        if self._codec is None:
            from ..codec import default_codec  # raises GpuCasUnavailable: no CPU fallback

            self._codec = default_codec()
        return self._codec

    def set_blob_cache(self, blob_cache):
        self._blob_cache = blob_cache

    def _session(self, codec):
        """Exclusive use of the codec's page-locked arenas while views into them are alive (see GpuCodec.session)."""
        session = getattr(codec, "session", None)
        return session() if session else nullcontext()

    _SMALL = 4096          # blobs up to this size are deduplicated by content on the host before any GPU work
    _SMALL_CACHE = 4096    # entries kept

    def _remember_small(self, content, key):
        cache = self._small_keys
        cache[content] = key
        if len(cache) > self._SMALL_CACHE:
            cache.popitem(last=False)

    def _path_for(self, key):
        return self._storage_impl.path_join(self._prefix, key[:2], key)

    def _write(self, items):
        """items: [(path, bytes-like, meta)] -> storage (zero-copy when the backend can)."""
        if not items:
            return
        if hasattr(self._storage_impl, "save_views"):
            self._storage_impl.save_views(iter(items))
        else:
            self._storage_impl.save_bytes(
                ((path, (BytesIO(bytes(view)), meta)) for path, view, meta in items),
                overwrite=True,
                len_hint=len(items),
            )

    def _windows(self, blob_iter):
        batch, size = [], 0
        for blob in blob_iter:
            batch.append(blob)
            size += len(blob[1]) if isinstance(blob, tuple) else len(blob)
            if size >= self._window_bytes:
                yield batch
                batch, size = [], 0
        if batch:
            yield batch

    # -- save --------------------------------------------------------------------------
    def save_blobs(self, blob_iter, raw=False, len_hint=0, is_transfer=False):
        """Iterator[bytes] -> [save_blobs_result(uri, key)] in input order (uri is None unless raw).

        is_transfer=True: items are (key, bytes, meta) triples coming out of another store's
        load_blobs(is_transfer=True); they are stored verbatim as raw blobs
        (content_addressed_store.py:82-96)."""
        results = []
        storage = self._storage_impl
        for window in self._windows(blob_iter):
            to_write = []
            if is_transfer:
                for key, blob_data, _meta in window:
                    path = self._path_for(key)
                    results.append(self.save_blobs_result(uri=storage.full_uri(path), key=key))
                    to_write.append((path, blob_data, {"cas_raw": True, "cas_version": 1}))
                self._write(to_write)
                continue
            # ---- dedup BEFORE pack (content_addressed_store.py:98-107 hashes, then skips what exists): identical
            # small blobs of the window collapse to one, blobs keyed earlier by this store need no GPU work at all,
            # and only what is left goes to the codec.  Large blobs are hashed and packed in the same GPU pass.
            n = len(window)
            keys = [None] * n
            cached, dup_of, first_of, todo = set(), {}, {}, []
            for i, blob in enumerate(window):
                if len(blob) <= self._SMALL:
                    content = bytes(blob)
                    known = self._small_keys.get(content)
                    if known is not None:
                        keys[i] = known
                        cached.add(i)
                        continue
                    j = first_of.setdefault(content, i)
                    if j != i:
                        dup_of[i] = j
                        continue
                todo.append(i)
            meta = {"cas_raw": raw, "cas_version": 1}
            codec = self.codec if todo else None
            lost = []
            with self._session(codec):
                payload_of = {}
                if todo:
                    sub = [window[i] for i in todo]
                    sub_keys, sub_payloads = (codec.hash_batch(sub), sub) if raw else codec.pack_batch(sub)
                    for i, k, pl in zip(todo, sub_keys, sub_payloads):
                        keys[i], payload_of[i] = k, pl
                    for content, i in first_of.items():
                        self._remember_small(content, keys[i])
                for i, j in dup_of.items():
                    keys[i] = keys[j]
                paths = [self._path_for(k) for k in keys]
                present = storage.is_file(paths)  # existing keys are never rewritten (:107)
                seen = set()
                for i in range(n):
                    results.append(self.save_blobs_result(uri=storage.full_uri(paths[i]) if raw else None, key=keys[i]))
                    if present[i] or keys[i] in seen:
                        continue
                    seen.add(keys[i])
                    if i in payload_of:
                        to_write.append((paths[i], payload_of[i], meta))
                    elif i in cached:
                        lost.append(i)                   # keyed from the cache, but its file is not there (any more)
                self._write(to_write)
            if lost:
                codec = self.codec
                with self._session(codec):
                    sub = [window[i] for i in lost]
                    payloads = sub if raw else codec.pack_batch(sub)[1]
                    self._write([(paths[i], pl, meta) for i, pl in zip(lost, payloads)])
            trim = getattr(codec, "trim", None) if codec is not None else None
            if trim:
                trim()
        return results

    # -- load --------------------------------------------------------------------------
    def load_blobs(self, keys, force_raw=False, is_transfer=False):
        """Iterable[key] -> iterator of (key, bytes) [(key, bytes, meta) when is_transfer]; order
        is not guaranteed; one result per requested path (content_addressed_store.py:122-204)."""
        for item in self._load(keys, force_raw, is_transfer, as_bytes=True):
            yield item

    def _load(self, keys, force_raw, is_transfer, as_bytes):
        for batch in self._load_batches(keys, force_raw, is_transfer, as_bytes):
            for item in batch:
                yield item

    def _load_batches(self, keys, force_raw, is_transfer, as_bytes):
        """Like load_blobs, but yields LISTS of results: one list per decoded GPU window (with as_bytes=False its
        blobs are views over the codec's page-locked output, valid until the generator is advanced), singletons
        for everything else.  TaskDataStore.load_artifacts unpickles a window at a time from it."""
        storage = self._storage_impl
        pending = []
        for key in keys:
            blob = self._blob_cache.load_key(key) if self._blob_cache else None
            if blob is not None:
                yield [(key, blob, {"cas_raw": False, "cas_version": 1}) if is_transfer else (key, blob)]
            else:
                pending.append(self._path_for(key))
        if not pending:
            return
        per_blob_unpack = "_unpack_v1" in self.__dict__ or type(self)._unpack_v1 is not ContentAddressedStore._unpack_v1
        with storage.load_bytes(pending) as loaded:
            # (key, path_key, meta, file_path) of blobs to decode on the GPU.  A window is bounded by its packed
            # bytes, by its DECODED bytes (a constant array inflates 1000x: the page-locked output arena and the
            # unpickled objects scale with that, not with the file sizes) and by its blob count.
            window, packed_bytes, unpacked_bytes = [], 0, 0
            for path_key, file_path, meta in loaded:
                key = storage.path_split(path_key)[-1]
                if force_raw or (meta and meta.get("cas_raw", False)):
                    with open(file_path, "rb") as f:
                        blob = f.read()
                    if self._blob_cache:
                        self._blob_cache.store_key(key, blob)
                    yield [(key, blob, meta) if is_transfer else (key, blob)]
                    continue
                if meta is None:
                    unpack_code = self._unpack_backward_compatible  # pre-meta datastores
                    version = 1
                else:
                    version = meta.get("cas_version", -1)
                    if version == -1:
                        raise DataException("Could not extract encoding version for '%s'" % path_key)
                    unpack_code = getattr(self, "_unpack_v%d" % version, None)
                    if unpack_code is None:
                        raise DataException(
                            "Unknown encoding version %d for '%s' -- the artifact is either corrupt or you "
                            "need to update Metaflow to the latest version" % (version, path_key)
                        )
                if per_blob_unpack or version != 1:
                    with open(file_path, "rb") as f:
                        try:
                            blob = unpack_code(f)
                        except Exception as e:
                            raise DataException("Could not unpack artifact '%s': %s" % (path_key, e))
                    if self._blob_cache:
                        self._blob_cache.store_key(key, blob)
                    yield [(key, blob, meta) if is_transfer else (key, blob)]
                    continue
                size = os.path.getsize(file_path)
                hint = packed_size_hint(file_path, size)
                if window and (unpacked_bytes + hint > self._window_unpacked or len(window) >= self._window_blobs):
                    for out in self._decode_window(window, is_transfer, as_bytes):
                        yield out
                    window, packed_bytes, unpacked_bytes = [], 0, 0
                window.append((key, path_key, meta, file_path))
                packed_bytes += size
                unpacked_bytes += hint
                if packed_bytes >= self._window_bytes:
                    for out in self._decode_window(window, is_transfer, as_bytes):
                        yield out
                    window, packed_bytes, unpacked_bytes = [], 0, 0
            if window:
                for out in self._decode_window(window, is_transfer, as_bytes):
                    yield out

    def _decode_window(self, window, is_transfer, as_bytes):
        """One GPU batch.  The codec is resolved only here (raw loads never need one), and its arenas stay
        locked until the consumer has advanced past the yielded list (its blobs may be views into them)."""
        codec = self.codec
        read_files = self._file_sources if getattr(codec, "accepts_file_sources", False) \
            else getattr(self._storage_impl, "read_files", None)
        with self._session(codec):
            yield list(self._flush_unpack(self._read_window(window, read_files), is_transfer, as_bytes))
        trim = getattr(codec, "trim", None)
        if trim:
            trim()

    @staticmethod
    def _file_sources(paths):
        from ..codec import FileSource

        return [FileSource(p, os.path.getsize(p)) for p in paths]

    @staticmethod
    def _read_window(window, read_files):
        """-> [(key, path_key, meta, packed bytes)]; the files of a window are read in parallel when the
        storage backend offers it."""
        if not window:
            return []
        paths = [w[3] for w in window]
        if read_files is not None:
            datas = read_files(paths)
        else:
            datas = []
            for p in paths:
                with open(p, "rb") as f:
                    datas.append(f.read())
        return [(k, pk, m, d) for (k, pk, m, _), d in zip(window, datas)]

    def _flush_unpack(self, batch, is_transfer, as_bytes):
        if not batch:
            return
        decoded = self.codec.unpack_batch([b[3] for b in batch], as_bytes=as_bytes or bool(self._blob_cache))
        for (key, path_key, meta, _), blob in zip(batch, decoded):
            if isinstance(blob, Exception):
                raise DataException("Could not unpack artifact '%s': %s" % (path_key, blob))
            if self._blob_cache:
                self._blob_cache.store_key(key, blob)
            yield (key, blob, meta) if is_transfer else (key, blob)

    # -- the versioned codec hooks of the reference, kept as single-blob entry points ---
    def _unpack_backward_compatible(self, blob):
        return self._unpack_v1(blob)

    def _pack_v1(self, blob):
        """bytes -> BytesIO of the RFC 1952 file (one-blob batch through the GPU codec)."""
        _, packed = self.codec.pack_batch([blob])
        return BytesIO(bytes(packed[0]))

    def _unpack_v1(self, blob):
        """file object -> decoded bytes (one-blob batch through the GPU codec)."""
        out = self.codec.unpack_batch([blob.read()])[0]
        if isinstance(out, Exception):
            raise out
        return out


class BlobCache(object):
    def load_key(self, key):
        pass

    def store_key(self, key, blob):
        pass
