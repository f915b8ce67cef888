"""Client-side file cache in front of the GPU datastore.

Same contract as metaflow/client/filecache.py: `FileCache` (:44-470) keeps decoded blobs and task
metadata under CLIENT_CACHE_PATH so that a second read of an artifact never touches the datastore,
bounded by CLIENT_CACHE_MAX_SIZE megabytes with oldest-first eviction that spares files younger than
NEW_FILE_QUARANTINE seconds (:333-346); `FileBlobCache` (:453-470) is the `BlobCache` handed to
`ContentAddressedStore.set_blob_cache` (cache file <cache>/<flow_ds_id>/<key[:2]>/<key>.blob, :459-463),
`TaskMetadataCache` (:405-450) the `MetadataCache` handed to `FlowDataStore.set_metadata_cache`
(<cache>/<task_ds_id>/<t[:2]>/<t>.cached with t = sha1("run/step/task/attempt/metadata")).

What is added for the GPU path: `get_artifacts_for_tasks` -- the artifacts of MANY tasks (a whole step of a
foreach, a join's inputs, `Run.data` over a step) are looked up in the cache first, and all misses of all
tasks go to the content-addressed store as ONE load (one batched GPU unpack) instead of one
`load_artifacts` per task as `DataArtifact.data` does (metaflow/client/core.py:973).
"""
import json
import os
import pickle
import time
from collections import OrderedDict
from hashlib import sha1
from tempfile import NamedTemporaryFile
from urllib.parse import urlparse

from ..datastore import FlowDataStore
from ..datastore.content_addressed_store import BlobCache
from ..datastore.flow_datastore import MetadataCache
from ..exception import MetaflowException
from .. import metaflow_config
from ..plugins.datastores.local_storage import LocalStorage

NEW_FILE_QUARANTINE = 10
DATASTORES = [LocalStorage]


class FileCacheException(MetaflowException):
    headline = "File cache error"


def _silent_unlink(path):
    try:
        os.unlink(path)
    except OSError:
        pass


class _CacheIndex(object):
    """What is in the cache directory, oldest first, and how much it weighs.  Built lazily (walking the directory
    can be expensive), appended to on every write, trimmed to the size budget oldest-first -- but never a file
    younger than NEW_FILE_QUARANTINE seconds, which another process may just be about to read."""

    SUFFIXES = (".cached", ".blob")

    def __init__(self, root, budget_bytes):
        self.root, self.budget, self.total, self.entries = root, budget_bytes, 0, None

    def _scan(self):
        found = []
        if os.path.isdir(self.root):
            for ds_id in os.listdir(self.root):
                for sub, _dirs, files in _two_levels(os.path.join(self.root, ds_id)):
                    for name in files:
                        if name.endswith(self.SUFFIXES):
                            path = os.path.join(sub, name)
                            try:
                                found.append((os.path.getctime(path), os.path.getsize(path), path))
                            except OSError:
                                pass
        found.sort()
        self.entries, self.total = found, sum(size for _, size, _ in found)

    def added(self, path):
        if self.entries is None:
            self._scan()
            if any(p == path for _, _, p in self.entries):
                return self.trim()
        size = os.path.getsize(path)
        self.entries.append((int(time.time()), size, path))
        self.total += size
        self.trim()

    def trim(self):
        now = time.time()
        while self.entries and self.total > self.budget and now - self.entries[0][0] >= NEW_FILE_QUARANTINE:
            _, size, path = self.entries.pop(0)
            self.total -= size
            try:
                os.remove(path)
            except OSError:
                pass                          # another client removed it already


def _two_levels(ds_dir):
    """<cache>/<ds id>/<two hex chars>/<files>: yields (subdir, [], files) for every second-level directory."""
    if not os.path.isdir(ds_dir):
        return
    for sub in os.listdir(ds_dir):
        subdir = os.path.join(ds_dir, sub)
        if os.path.isdir(subdir):
            yield subdir, [], os.listdir(subdir)


class FileCache(object):
    def __init__(self, cache_dir=None, max_size=None, codec=None):
        self._cache_dir = cache_dir if cache_dir is not None else metaflow_config.CLIENT_CACHE_PATH
        self._max_size = int(max_size if max_size is not None else metaflow_config.CLIENT_CACHE_MAX_SIZE)
        self._codec = codec                  # GpuCodec to hand to the datastores (None: the process-wide one)
        self._index = _CacheIndex(self._cache_dir, self._max_size * 1024 ** 2)
        self._blob_caches = {}               # flow_ds_id -> (FileBlobCache, TaskMetadataCache)
        self._store_caches = OrderedDict()   # flow_ds_id -> FlowDataStore, least recently used first

    @property
    def cache_dir(self):
        return self._cache_dir

    # ---------------------------------------------------------------- artifacts
    def get_artifact_size(self, ds_type, ds_root, attempt, flow_name, run_id, step_name, task_id, name):
        task_ds = self._get_task_datastore(ds_type, ds_root, flow_name, run_id, step_name, task_id, attempt)
        _, size = next(task_ds.get_artifact_sizes([name]))
        return size

    def get_artifact_by_location(self, ds_type, location, data_metadata, flow_name, run_id, step_name, task_id, name):
        ds_cls = self._get_datastore_storage_impl(ds_type)
        ds_root = ds_cls.get_datastore_root_from_location(location, flow_name)
        return self.get_artifact(ds_type, ds_root, data_metadata, flow_name, run_id, step_name, task_id, name)

    def get_artifact(self, ds_type, ds_root, data_metadata, flow_name, run_id, step_name, task_id, name):
        _, obj = next(self.get_artifacts(ds_type, ds_root, data_metadata, flow_name, run_id, step_name, task_id, [name]))
        return obj

    def get_all_artifacts(self, ds_type, ds_root, data_metadata, flow_name, run_id, step_name, task_id):
        ds = self._get_flow_datastore(ds_type, ds_root, flow_name)
        task_ds = ds.get_task_datastore(run_id, step_name, task_id, data_metadata=data_metadata)
        return task_ds.load_artifacts([n for n, _ in task_ds.items()])

    def get_artifacts(self, ds_type, ds_root, data_metadata, flow_name, run_id, step_name, task_id, names):
        ds = self._get_flow_datastore(ds_type, ds_root, flow_name)
        task_ds = ds.get_task_datastore(run_id, step_name, task_id, data_metadata=data_metadata)
        # load_artifacts goes through flow_datastore.ca_store, i.e. through this cache's FileBlobCache
        return task_ds.load_artifacts(names)

    def get_artifacts_for_tasks(self, ds_type, ds_root, flow_name, tasks, names):
        """tasks: [(run_id, step_name, task_id, data_metadata or None)] -> {(run_id, step_name, task_id): {name: obj}}.
        One content-addressed load for all tasks: cache hits are read from the cache files, every miss of
        every task is decoded in one GPU batch (and stored in the cache)."""
        ds = self._get_flow_datastore(ds_type, ds_root, flow_name)
        stores = [(t[:3], ds.get_task_datastore(t[0], t[1], t[2], data_metadata=t[3])) for t in tasks]
        wanted = {}                                           # key -> [(task id, name)]
        for tid, task_ds in stores:
            for name in names:
                if name in task_ds:
                    wanted.setdefault(task_ds._objects[name], []).append((tid, name))
        out = {tid: {} for tid, _ in stores}
        for key, blob in ds.ca_store.load_blobs(wanted.keys()):
            for tid, name in wanted[key]:
                out[tid][name] = pickle.loads(blob)           # every (task, name) gets its own object
        return out

    # ---------------------------------------------------------------- cache files
    def create_file(self, path, value):
        dirname = os.path.dirname(path)
        try:
            os.makedirs(dirname, exist_ok=True)
        except OSError:
            raise FileCacheException("Could not create directory: %s" % dirname)
        with NamedTemporaryFile(dir=dirname, prefix="dlobj", delete=False) as tmp:
            tmp_name = tmp.name
            try:
                tmp.write(value)
            except BaseException:
                tmp.close()
                _silent_unlink(tmp_name)
                raise
        try:
            os.rename(tmp_name, path)          # readers see the whole file or none of it
        except BaseException:
            _silent_unlink(tmp_name)
            raise
        self._index.added(path)

    def read_file(self, path):
        if os.path.exists(path):
            try:
                with open(path, "rb") as f:
                    return f.read()
            except IOError:
                pass                          # garbage collected by another process meanwhile
        return None

    # ---------------------------------------------------------------- ids
    @staticmethod
    def flow_ds_id(ds_type, ds_root, flow_name):
        p = urlparse(ds_root)
        return ".".join([ds_type, (p.netloc + p.path).replace("/", "_"), flow_name])

    @staticmethod
    def task_ds_id(ds_type, ds_root, flow_name, run_id, step_name, task_id, attempt):
        p = urlparse(ds_root)
        return ".".join([ds_type, (p.netloc + p.path).replace("/", "_"), flow_name, run_id, step_name, task_id, str(attempt)])

    # ---------------------------------------------------------------- datastores
    @staticmethod
    def _get_datastore_storage_impl(ds_type):
        impl = [d for d in DATASTORES if d.TYPE == ds_type]
        if not impl:
            raise FileCacheException("Datastore %s was not found" % ds_type)
        return impl[0]

    def _get_flow_datastore(self, ds_type, ds_root, flow_name):
        cache_id = self.flow_ds_id(ds_type, ds_root, flow_name)
        cached = self._store_caches.get(cache_id)
        if cached:
            self._store_caches.move_to_end(cache_id)
            return cached
        storage_impl = self._get_datastore_storage_impl(ds_type)
        fds = FlowDataStore(flow_name=flow_name, environment=None, storage_impl=storage_impl, ds_root=ds_root, codec=self._codec)
        caches = self._blob_caches.setdefault(
            cache_id, (FileBlobCache(self, cache_id), TaskMetadataCache(self, ds_type, ds_root, flow_name)))
        fds.ca_store.set_blob_cache(caches[0])
        fds.set_metadata_cache(caches[1])
        self._store_caches[cache_id] = fds
        if len(self._store_caches) > int(metaflow_config.CLIENT_CACHE_MAX_FLOWDATASTORE_COUNT):
            evicted, _ = self._store_caches.popitem(last=False)
            del self._blob_caches[evicted]
        return fds

    def _get_task_datastore(self, ds_type, ds_root, flow_name, run_id, step_name, task_id, attempt):
        return self._get_flow_datastore(ds_type, ds_root, flow_name).get_task_datastore(run_id, step_name, task_id, attempt=attempt)


class TaskMetadataCache(MetadataCache):
    def __init__(self, filecache, ds_type, ds_root, flow_name):
        self._filecache = filecache
        self._ds_type = ds_type
        self._ds_root = ds_root
        self._flow_name = flow_name

    def _path(self, run_id, step_name, task_id, attempt):
        if attempt is None:
            raise MetaflowException("Attempt number must be specified to use task metadata cache.")
        cache_id = self._filecache.task_ds_id(self._ds_type, self._ds_root, self._flow_name, run_id, step_name, task_id, attempt)
        token = "%s.cached" % sha1(os.path.join(run_id, step_name, task_id, str(attempt), "metadata").encode("utf-8")).hexdigest()
        return os.path.join(self._filecache.cache_dir, cache_id, token[:2], token)

    def load_metadata(self, run_id, step_name, task_id, attempt):
        d = self._filecache.read_file(self._path(run_id, step_name, task_id, attempt))
        if d:
            return json.loads(d)

    def store_metadata(self, run_id, step_name, task_id, attempt, metadata_dict):
        self._filecache.create_file(self._path(run_id, step_name, task_id, attempt), json.dumps(metadata_dict).encode("utf-8"))


class FileBlobCache(BlobCache):
    def __init__(self, filecache, cache_id):
        self._filecache = filecache
        self._cache_id = cache_id

    def _path(self, key):
        return os.path.join(self._filecache.cache_dir, self._cache_id, key[:2], "%s.blob" % key)

    def load_key(self, key):
        return self._filecache.read_file(self._path(key))

    def store_key(self, key, blob):
        self._filecache.create_file(self._path(key), blob)
