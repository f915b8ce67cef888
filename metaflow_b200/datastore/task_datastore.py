"""Per-task datastore: artifacts (through the content-addressed store), JSON metadata files, logs.

Keeps the surface and every on-disk byte of metaflow/datastore/task_datastore.py:58-1051:
`<attempt>.attempt.json`, `<attempt>.data.json` ({"datastore","version":"1.0","attempt",
"python_version","objects","info"}), `<attempt>.DONE.lock`, `_info[name] = {"size","type",
"encoding":"gzip+pickle-v4"}` (:371-375), mode guards (:23-50), latest-done-attempt resolution
(:133-182), clone / passdown_partial (:714-740), persist (:744-809).

The hot path differs in shape, not in result: save_artifacts (:329-383) pickles the whole artifact
set and hands it to ContentAddressedStore.save_blobs as ONE batch (SHA-1 + DEFLATE + CRC-32 on the
GPUs); load_artifacts (:386-442) loads all requested keys as one batch and unpickles straight out
of the page-locked output arena.
"""
import json
import os
import pickle
import sys
import time
from collections import defaultdict
from functools import wraps
from io import BufferedIOBase, FileIO, RawIOBase
from types import FunctionType, MethodType

from .. import metaflow_config
from ..exception import DataException, MetaflowInternalError, UnpicklableArtifactException
from ..metadata_provider import DataArtifact, MetaDatum
from ..parameters import Parameter
from ..serializer import dumps_segments, loads_batch
from ..util import Path, is_stringish, to_fileobj

_included_file_type = "<class 'metaflow.includefile.IncludedFile'>"


def only_if_not_done(f):
    @wraps(f)
    def guarded(self, *args, **kwargs):
        if self._is_done_set:
            raise MetaflowInternalError(
                "Tried to write to datastore (method %s) after it was marked .done()" % f.__name__
            )
        return f(self, *args, **kwargs)

    return guarded


def require_mode(mode):
    def deco(f):
        @wraps(f)
        def guarded(self, *args, **kwargs):
            if mode is not None and self._mode != mode:
                raise MetaflowInternalError(
                    "Attempting a datastore operation '%s' requiring mode '%s' but have mode '%s'"
                    % (f.__name__, mode, self._mode)
                )
            return f(self, *args, **kwargs)

        return guarded

    return deco


class ArtifactTooLarge(object):
    def __str__(self):
        return "< artifact too large >"


_pool = None


def _unpickle_pool():
    global _pool
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor

        _pool = ThreadPoolExecutor(max_workers=max(2, min(16, len(os.sched_getaffinity(0)))), thread_name_prefix="mf-unpickle")
    return _pool


def _read_only(attr):
    """A property that exposes a private attribute."""
    return property(lambda self: getattr(self, attr))


class TaskDataStore(object):
    METADATA_ATTEMPT_SUFFIX = "attempt.json"
    METADATA_DONE_SUFFIX = "DONE.lock"
    METADATA_DATA_SUFFIX = "data.json"

    @staticmethod
    def metadata_name_for_attempt(name, attempt):
        return name if attempt is None else "%d.%s" % (attempt, name)

    @staticmethod
    def parse_attempt_metadata(name):
        return name.split(".", 1)

    def __init__(self, flow_datastore, run_id, step_name, task_id, attempt=None, data_metadata=None,
                 mode="r", allow_not_done=False, persist=True):
        self._storage_impl = flow_datastore._storage_impl
        self.TYPE = self._storage_impl.TYPE
        self._ca_store = flow_datastore.ca_store
        self._environment = flow_datastore.environment
        self._run_id, self._step_name, self._task_id = run_id, step_name, task_id
        self._path = self._storage_impl.path_join(flow_datastore.flow_name, run_id, step_name, task_id)
        self._mode = mode
        self._attempt = attempt
        self._metadata = flow_datastore.metadata
        self._parent = flow_datastore
        self._persist = persist
        self._encodings = {"pickle-v2", "gzip+pickle-v2", "pickle-v4", "gzip+pickle-v4"}
        self._is_done_set = False
        self._objects, self._info = {}, {}

        if mode == "w":
            return
        if mode == "r":
            if data_metadata is not None:
                self._objects = data_metadata.get("objects", {})
                self._info = data_metadata.get("info", {})
                return
            latest = self._latest_started_attempt(stop_at_gap=True)
            if self._attempt is None:
                self._attempt = latest
            elif latest is None or self._attempt > latest:
                return  # that attempt never started: nothing to load
            # only a finished attempt exposes data, unless the caller accepts unfinished ones
            if self.has_metadata(self.METADATA_DONE_SUFFIX):
                data_obj = self.load_metadata([self.METADATA_DATA_SUFFIX])[self.METADATA_DATA_SUFFIX]
                if data_obj is not None:
                    self._objects = data_obj.get("objects", {})
                    self._info = data_obj.get("info", {})
            elif self._attempt is None or not allow_not_done:
                raise DataException("No completed attempts of the task was found for task '%s'" % self._path)
            return
        if mode == "d":
            if self._attempt is None:
                self._attempt = self._latest_started_attempt(stop_at_gap=False)
            if not allow_not_done and not self.has_metadata(self.METADATA_DONE_SUFFIX):
                raise DataException("No completed attempts of the task was found for task '%s'" % self._path)
            return
        raise DataException("Unknown datastore mode: '%s'" % self._mode)

    def _latest_started_attempt(self, stop_at_gap):
        latest = None
        for i in range(metaflow_config.MAX_ATTEMPTS):
            name = self.metadata_name_for_attempt(self.METADATA_ATTEMPT_SUFFIX, i)
            if self.has_metadata(name, add_attempt=False):
                latest = i
            elif stop_at_gap and latest is not None:
                break
        return latest

    # -- identity --------------------------------------------------------------------------
    run_id = _read_only("_run_id")
    step_name = _read_only("_step_name")
    task_id = _read_only("_task_id")
    attempt = _read_only("_attempt")

    @property
    def pathspec(self):
        return "%s/%s/%s" % (self._run_id, self._step_name, self._task_id)

    @property
    def ds_metadata(self):
        """What a MetadataCache stores for this task: copies of the name->key and name->info maps."""
        return {"objects": dict(self._objects), "info": dict(self._info)}

    @property
    def pathspec_index(self):
        idx = ",".join(str(f.index) for f in self["_foreach_stack"])
        if "_iteration_stack" in self:
            itr = ",".join(str(f) for f in self["_iteration_stack"])
            return "%s/%s[%s][%s]" % (self._run_id, self._step_name, idx, itr)
        return "%s/%s[%s]" % (self._run_id, self._step_name, idx)

    @property
    def parent_datastore(self):
        return self._parent

    @require_mode(None)
    def get_log_location(self, logprefix, stream):
        name = self._metadata_name_for_attempt(self._get_log_location(logprefix, stream))
        return self._storage_impl.full_uri(self._storage_impl.path_join(self._path, name))

    @require_mode("r")
    def keys_for_artifacts(self, names):
        return [self._objects.get(name) for name in names]

    # -- lifecycle --------------------------------------------------------------------------
    @only_if_not_done
    @require_mode("w")
    def init_task(self):
        self.save_metadata({self.METADATA_ATTEMPT_SUFFIX: {"time": time.time()}})

    @only_if_not_done
    @require_mode("w")
    def transfer_artifacts(self, other_datastore, names=None):
        """Bring blobs over from a datastore with a different root (spin); task_datastore.py:263-325."""
        if (other_datastore.TYPE == self.TYPE
                and other_datastore._storage_impl.datastore_root == self._storage_impl.datastore_root):
            return
        wanted = list(other_datastore._objects) if names is None else [n for n in names if n in other_datastore._objects]
        if not wanted:
            return
        cas = self._ca_store
        missing = []
        for sha in (other_datastore._objects[n] for n in wanted):
            path = cas._storage_impl.path_join(cas._prefix, sha[:2], sha)
            if not cas._storage_impl.is_file([path])[0]:
                missing.append(sha)
        if missing:
            cas.save_blobs(other_datastore._ca_store.load_blobs(missing, is_transfer=True), is_transfer=True)

    # -- the hot path -----------------------------------------------------------------------
    @only_if_not_done
    @require_mode("w")
    def save_artifacts(self, artifacts_iter, len_hint=0):
        """Iterator[(name, object)] -> pickled (protocol 4), keyed, packed and stored as one batch."""
        names = []

        def blobs():
            for name, obj in artifacts_iter:
                try:
                    blob = dumps_segments(obj)         # the bytes of pickle.dumps(obj, protocol=4), arrays by reference
                except TypeError as e:
                    raise UnpicklableArtifactException(name) from e
                self._info[name] = {"size": len(blob), "type": str(type(obj)), "encoding": "gzip+pickle-v4"}
                names.append(name)
                yield blob

        for name, res in zip(names, self._ca_store.save_blobs(blobs(), len_hint=len_hint)):
            self._objects[name] = res.key

    @require_mode(None)
    def load_artifacts(self, names):
        """[name] -> iterator of (name, object); every name gets its own freshly unpickled object."""
        if not self._info:
            raise DataException(
                "Datastore for task '%s' does not have the required metadata to load artifacts" % self._path
            )
        by_key = defaultdict(list)
        for name in names:
            info = self._info.get(name)
            encoding = info.get("encoding", "gzip+pickle-v2") if info else "gzip+pickle-v2"
            if encoding not in self._encodings:
                raise DataException("Python 3.6 or later is required to load artifact '%s'" % name)
            by_key[self._objects[name]].append(name)
        loader = getattr(self._ca_store, "_load_batches", None)
        if loader is None:
            for key, blob in self._ca_store.load_blobs(by_key.keys()):
                for name in by_key[key]:
                    yield name, pickle.loads(blob)
            return
        # one decoded GPU window at a time: the blobs are views over the codec's page-locked output, and
        # the whole window is unpickled before the generator moves on (payload copies on a thread pool)
        for batch in loader(by_key.keys(), False, False, as_bytes=False):
            names, blobs = [], []
            for key, blob in batch:
                for name in by_key[key]:          # every name gets its own object (task_datastore.py:439-442)
                    names.append(name)
                    blobs.append(blob)
            big = sum(len(b) for b in blobs) >= (8 << 20) and len(blobs) > 1
            for name, obj in zip(names, loads_batch(blobs, _unpickle_pool() if big else None)):
                yield name, obj

    @require_mode("r")
    def get_artifact_sizes(self, names):
        for name in names:
            info = self._info.get(name)
            yield name, (self[name].size if info["type"] == _included_file_type else info.get("size", 0))

    # -- metadata files -----------------------------------------------------------------------
    @only_if_not_done
    @require_mode("w")
    def save_metadata(self, contents, allow_overwrite=True, add_attempt=True):
        return self._save_file({k: json.dumps(v).encode("utf-8") for k, v in contents.items()}, allow_overwrite, add_attempt)

    @require_mode("w")
    def _dangerous_save_metadata_post_done(self, contents, allow_overwrite=True, add_attempt=True):
        return self._save_file({k: json.dumps(v).encode("utf-8") for k, v in contents.items()}, allow_overwrite, add_attempt)

    @require_mode("r")
    def load_metadata(self, names, add_attempt=True):
        return {k: (json.loads(v) if v is not None else None) for k, v in self._load_file(names, add_attempt).items()}

    @require_mode(None)
    def has_metadata(self, name, add_attempt=True):
        leaf = self._metadata_name_for_attempt(name) if add_attempt else name
        return self._storage_impl.is_file([self._storage_impl.path_join(self._path, leaf)])[0]

    @require_mode(None)
    def get(self, name, default=None):
        if self._objects and name in self._objects:
            try:
                return self[name]
            except DataException:
                return default
        return default

    @require_mode("r")
    def is_none(self, name):
        if not self._info:
            return True
        info = self._info.get(name)
        if info and info.get("type") == str(type(None)):
            return True
        return self.get(name) is None

    @only_if_not_done
    @require_mode("w")
    def done(self):
        self.save_metadata({
            self.METADATA_DATA_SUFFIX: {
                "datastore": self.TYPE,
                "version": "1.0",
                "attempt": self._attempt,
                "python_version": sys.version,
                "objects": self._objects,
                "info": self._info,
            },
            self.METADATA_DONE_SUFFIX: "",
        })
        if self._metadata:
            self._metadata.register_metadata(
                self._run_id, self._step_name, self._task_id,
                [MetaDatum(field="attempt-done", value=str(self._attempt), type="attempt-done",
                           tags=["attempt_id:{0}".format(self._attempt)])],
            )
            self._metadata.register_data_artifacts(
                self.run_id, self.step_name, self.task_id, self._attempt,
                [DataArtifact(name=var, ds_type=self.TYPE, ds_root=self._storage_impl.datastore_root, url=None,
                              sha=sha, type=self._info[var]["encoding"]) for var, sha in self._objects.items()],
            )
        self._is_done_set = True

    @only_if_not_done
    @require_mode("w")
    def clone(self, origin):
        self._objects = origin._objects
        self._info = origin._info

    @only_if_not_done
    @require_mode("w")
    def passdown_partial(self, origin, variables):
        for var in variables:
            sha = origin._objects.get(var)
            if sha:
                self._objects[var] = sha
                self._info[var] = origin._info[var]

    @only_if_not_done
    @require_mode("w")
    def persist(self, flow):
        """Save every artifact attribute of `flow`.  DESTRUCTIVE: user artifacts are removed from
        the flow object as they are handed over (task_datastore.py:744-809)."""
        if not self._persist:
            return
        if flow._datastore:
            self._objects.update(flow._datastore._objects)
            self._info.update(flow._datastore._info)
        found = []
        for var in dir(flow):
            if var.startswith("__") or var in flow._EPHEMERAL:
                continue
            if isinstance(getattr(flow.__class__, var, None), property):
                continue
            val = getattr(flow, var)
            if isinstance(val, (MethodType, FunctionType, Parameter)):
                continue
            found.append((var, val))
        if hasattr(flow._datastore, "orig_datastore"):
            untouched = set(flow._datastore._objects) - {v for v, _ in found}
            if untouched:
                self.transfer_artifacts(flow._datastore.orig_datastore, names=list(untouched))

        def handover():
            while found:
                var, val = found.pop()
                if not var.startswith("_") and var != "name":
                    delattr(flow, var)
                yield var, val

        self.save_artifacts(handover(), len_hint=len(found))

    # -- logs --------------------------------------------------------------------------------
    @only_if_not_done
    @require_mode("w")
    def save_logs(self, logsource, stream_data):
        files = {}
        for stream, data in stream_data.items():
            name = self._get_log_location(logsource, stream)
            files[name] = FileIO(str(data), mode="r") if isinstance(data, Path) else data
        self._save_file(files)

    @require_mode("d")
    def scrub_logs(self, logsources, stream, attempt_override=None):
        names = {self._metadata_name_for_attempt(self._get_log_location(s, stream), attempt_override=attempt_override): s
                 for s in logsources}
        names[self._metadata_name_for_attempt("%s.log" % stream, attempt_override)] = stream
        present = [p for p in names if self.has_metadata(p, add_attempt=False)]
        self._save_file({p: bytes("[REDACTED %s %s]" % (names[p], stream), "utf-8") for p in present},
                        add_attempt=False, allow_overwrite=True)

    @require_mode("r")
    def load_log_legacy(self, stream, attempt_override=None):
        name = self._metadata_name_for_attempt("%s.log" % stream, attempt_override)
        r = self._load_file([name], add_attempt=False)[name]
        return r if r is not None else b""

    @require_mode("r")
    def load_logs(self, logsources, stream, attempt_override=None):
        names = {self._metadata_name_for_attempt(self._get_log_location(s, stream), attempt_override=attempt_override): s
                 for s in logsources}
        loaded = self._load_file(names.keys(), add_attempt=False)
        return [(names[k], v if v is not None else b"") for k, v in loaded.items()]

    @require_mode("r")
    def get_legacy_log_size(self, stream):
        name = self._metadata_name_for_attempt("%s.log" % stream)
        return self._storage_impl.size_file(self._storage_impl.path_join(self._path, name))

    @require_mode("r")
    def get_log_size(self, logsources, stream):
        total = 0
        for s in logsources:
            leaf = self._metadata_name_for_attempt(self._get_log_location(s, stream))
            size = self._storage_impl.size_file(self._storage_impl.path_join(self._path, leaf))
            total += size or 0
        return total

    # -- dict-like access ---------------------------------------------------------------------
    @require_mode(None)
    def items(self):
        return self._objects.items() if self._objects else {}

    @require_mode(None)
    def to_dict(self, show_private=False, max_value_size=None, include=None):
        out = {}
        for k, _ in self.items():
            if (include and k not in include) or (k[0] == "_" and not show_private):
                continue
            info = self._info[k]
            included = info["type"] == _included_file_type
            if max_value_size is not None:
                size = self[k].size if included else info.get("size", 0)
                if size == 0 or size > max_value_size:
                    out[k] = ArtifactTooLarge()
                    continue
            out[k] = self[k].decode(k) if included else self[k]
        return out

    @require_mode("r")
    def format(self, **kwargs):
        lines = []
        for k, v in self.to_dict(**kwargs).items():
            size = self[k].size if self._info[k]["type"] == _included_file_type else self._info[k]["size"]
            lines.append((k, "*{key}* [size: {size} type: {type}] = {value}".format(
                key=k, value=v, size=size, type=self._info[k]["type"])))
        return "\n".join(line for _, line in sorted(lines))

    @require_mode(None)
    def __contains__(self, name):
        return bool(self._objects) and name in self._objects

    @require_mode(None)
    def __getitem__(self, name):
        _, obj = next(self.load_artifacts([name]))
        return obj

    @require_mode("r")
    def __iter__(self):
        return iter(self._objects) if self._objects else iter([])

    @require_mode("r")
    def __str__(self):
        return self.format(show_private=True, max_value_size=1000)

    # -- file helpers --------------------------------------------------------------------------
    def _metadata_name_for_attempt(self, name, attempt_override=None):
        return self.metadata_name_for_attempt(name, self._attempt if attempt_override is None else attempt_override)

    @staticmethod
    def _get_log_location(logprefix, stream):
        return "%s_%s.log" % (logprefix, stream)

    def _leaf_path(self, name, add_attempt):
        leaf = self._metadata_name_for_attempt(name) if add_attempt else name
        return self._storage_impl.path_join(self._path, leaf)

    def _save_file(self, contents, allow_overwrite=True, add_attempt=True):
        def files():
            for name, value in contents.items():
                path = self._leaf_path(name, add_attempt)
                if isinstance(value, (RawIOBase, BufferedIOBase)) and value.readable():
                    yield path, value
                elif is_stringish(value):
                    yield path, to_fileobj(value)
                else:
                    raise DataException(
                        "Metadata '%s' for task '%s' has an invalid type: %s" % (name, self._path, type(value))
                    )

        self._storage_impl.save_bytes(files(), overwrite=allow_overwrite)

    def _load_file(self, names, add_attempt=True):
        results = {}
        with self._storage_impl.load_bytes([self._leaf_path(n, add_attempt) for n in names]) as loaded:
            for key, path, _meta in loaded:
                base = self._storage_impl.basename(key)
                name = self.parse_attempt_metadata(base)[1] if add_attempt else base
                if path is None:
                    results[name] = None
                else:
                    with open(path, "rb") as f:
                        results[name] = f.read()
        return results
