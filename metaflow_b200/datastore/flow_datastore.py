"""Flow-level datastore: owns ONE content-addressed store (<flow>/data) and hands out TaskDataStores.

Surface of metaflow/datastore/flow_datastore.py:13-398: constructor (:16-67, CAS created at :65),
get_task_datastores (:79-254: latest attempt that started AND finished; pathspecs keep caller
order :244-254), get_task_datastore (:257-346), save_data / load_data for raw blobs (:348-388),
MetadataCache (:391-398).  A `codec` can be injected (tests inject the CPU oracle; production
uses the GPU codec and fails loudly without one).
"""
import json
from abc import ABC, abstractmethod

from .. import metaflow_config
from ..metaflow_profile import from_start
from .content_addressed_store import ContentAddressedStore
from .task_datastore import TaskDataStore


class FlowDataStore(object):
    default_storage_impl = None

    def __init__(self, flow_name, environment=None, metadata=None, event_logger=None, monitor=None,
                 storage_impl=None, ds_root=None, codec=None):
        storage_impl = storage_impl if storage_impl else self.default_storage_impl
        if storage_impl is None:
            raise RuntimeError("No datastore storage implementation specified")
        self._storage_impl = storage_impl(ds_root)
        self.TYPE = self._storage_impl.TYPE
        self.flow_name = flow_name
        self.environment = environment
        self.metadata = metadata
        self.logger = event_logger
        self.monitor = monitor
        self.ca_store = ContentAddressedStore(
            self._storage_impl.path_join(self.flow_name, "data"), self._storage_impl, codec=codec
        )
        self._metadata_cache = None

    @property
    def datastore_root(self):
        return self._storage_impl.datastore_root

    def set_metadata_cache(self, cache):
        self._metadata_cache = cache

    def _task_urls(self, run_id, steps, pathspecs):
        st = self._storage_impl
        if pathspecs:
            return [st.path_join(self.flow_name, ps) for ps in pathspecs]
        run_prefix = st.path_join(self.flow_name, run_id)
        if steps:
            step_urls = [st.path_join(run_prefix, s) for s in steps]
        else:
            step_urls = [e.path for e in st.list_content([run_prefix]) if e.is_file is False]
        return [e.path for e in st.list_content(step_urls) if e.is_file is False]

    def get_task_datastores(self, run_id=None, steps=None, pathspecs=None, allow_not_done=False, attempt=None,
                            include_prior=False, mode="r", join_type=None, orig_flow_datastore=None,
                            spin_artifacts=None):
        st = self._storage_impl
        attempts = range(metaflow_config.MAX_ATTEMPTS)
        if attempt is not None and attempt <= metaflow_config.MAX_ATTEMPTS - 1:
            attempts = range(attempt + 1) if include_prior else [attempt]
        suffixes = (TaskDataStore.METADATA_DATA_SUFFIX, TaskDataStore.METADATA_ATTEMPT_SUFFIX,
                    TaskDataStore.METADATA_DONE_SUFFIX)
        urls = []
        for task_url in self._task_urls(run_id, steps, pathspecs):
            parts = task_url.rstrip("/").split("/")
            task_attempts = [int(parts[4])] if len(parts) == 5 else attempts  # flow/run/step/task[/attempt]
            for a in task_attempts:
                for suffix in suffixes:
                    urls.append(st.path_join(task_url, TaskDataStore.metadata_name_for_attempt(suffix, a)))
        started, done, data_objs = {}, set(), {}
        with st.load_bytes(urls) as found:
            for key, path, _meta in found:
                if path is None:
                    continue
                _, run, step, task, fname = st.path_split(key)
                a, fname = TaskDataStore.parse_attempt_metadata(fname)
                a = int(a)
                if fname == TaskDataStore.METADATA_DONE_SUFFIX:
                    done.add((run, step, task, a))
                elif fname == TaskDataStore.METADATA_ATTEMPT_SUFFIX:
                    started[(run, step, task)] = max(started.get((run, step, task), 0), a)
                elif fname == TaskDataStore.METADATA_DATA_SUFFIX:
                    with open(path, encoding="utf-8") as f:
                        data_objs[(run, step, task, a)] = json.load(f)
        latest_started = set((r, s, t, a) for (r, s, t), a in started.items())
        if allow_not_done:
            chosen = done.union(latest_started) if include_prior else latest_started
        else:
            chosen = done if include_prior else (latest_started & done)
        chosen = list(chosen)
        if pathspecs:  # callers (foreach joins) rely on getting datastores back in pathspec order
            order = {(ps.split("/")[1], ps.split("/")[2]): i for i, ps in enumerate(pathspecs)}
            chosen.sort(key=lambda v: order[v[1], v[2]])
        return [
            self.get_task_datastore(r, s, t, a, data_objs.get((r, s, t, a)), mode, allow_not_done, join_type,
                                    orig_flow_datastore, spin_artifacts)
            for (r, s, t, a) in chosen
        ]

    def get_task_datastore(self, run_id, step_name, task_id, attempt=None, data_metadata=None, mode="r",
                           allow_not_done=False, join_type=None, orig_flow_datastore=None, spin_artifacts=None,
                           persist=True):
        if orig_flow_datastore is not None:
            from .spin_datastore import SpinTaskDatastore

            if join_type is not None:
                spin_artifacts = spin_artifacts.get("%s/%s/%s" % (run_id, step_name, task_id), {})
            orig = orig_flow_datastore.get_task_datastore(
                run_id, step_name, task_id, attempt=attempt, data_metadata=data_metadata, mode=mode,
                allow_not_done=allow_not_done, persist=persist)
            return SpinTaskDatastore(self.flow_name, run_id, step_name, task_id, orig, spin_artifacts)
        cacheable = self._metadata_cache is not None and attempt is not None and allow_not_done is False
        hit = False
        if cacheable and data_metadata is None:
            data_metadata = self._metadata_cache.load_metadata(run_id, step_name, task_id, attempt)
            hit = data_metadata is not None
        from_start("FlowDataStore: get_task_datastore for type %s %s metadata"
                   % (self.TYPE, "without" if data_metadata is None else "with"))
        ds = TaskDataStore(self, run_id, step_name, task_id, attempt=attempt, data_metadata=data_metadata,
                           mode=mode, allow_not_done=allow_not_done, persist=persist)
        if cacheable and not hit:
            self._metadata_cache.store_metadata(run_id, step_name, task_id, attempt, ds.ds_metadata)
        return ds

    def save_data(self, data_iter, len_hint=0):
        """Raw blobs (code packages, IncludeFile): key = GPU SHA-1, bytes stored verbatim -> [(uri, key)]."""
        return [(r.uri, r.key) for r in self.ca_store.save_blobs(data_iter, raw=True, len_hint=len_hint)]

    def load_data(self, keys, force_raw=False):
        for key, blob in self.ca_store.load_blobs(keys, force_raw=force_raw):
            yield key, blob


class MetadataCache(ABC):
    @abstractmethod
    def load_metadata(self, run_id, step_name, task_id, attempt):
        raise NotImplementedError()

    @abstractmethod
    def store_metadata(self, run_id, step_name, task_id, attempt, metadata_dict):
        raise NotImplementedError()
