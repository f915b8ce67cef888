"""Batch-open many task datastores and prefetch artifacts shared by a join into one blob cache.

metaflow/datastore/datastore_set.py:15-85.  The prefetch is the reference's only batched load
(used for join fan-in, task.py:252-279); here it is exactly one GpuCodec.unpack_batch over all
distinct keys, which is what turns a join's N serial gunzips into one GPU batch.
"""
from .content_addressed_store import BlobCache


class ImmutableBlobCache(BlobCache):
    """Serves only what was preloaded; never stores (datastore_set.py:76-85)."""

    def __init__(self, preloaded):
        self._preloaded = preloaded

    def load_key(self, key):
        return self._preloaded.get(key)

    def store_key(self, key, blob):
        pass


class TaskDataStoreSet(object):
    def __init__(self, flow_datastore, run_id, steps=None, pathspecs=None, prefetch_data_artifacts=None,
                 allow_not_done=False, join_type=None, orig_flow_datastore=None, spin_artifacts=None):
        self.task_datastores = flow_datastore.get_task_datastores(
            run_id, steps=steps, pathspecs=pathspecs, allow_not_done=allow_not_done, join_type=join_type,
            orig_flow_datastore=orig_flow_datastore, spin_artifacts=spin_artifacts)
        if prefetch_data_artifacts:
            wanted = set()
            for ds in self.task_datastores:
                wanted.update(ds.keys_for_artifacts(prefetch_data_artifacts))
            wanted.discard(None)
            preloaded = dict(flow_datastore.ca_store.load_blobs(wanted))   # ONE batched GPU unpack
            flow_datastore.ca_store.set_blob_cache(ImmutableBlobCache(preloaded))
        self.pathspec_index_cache = {}
        self.pathspec_cache = {}
        if not allow_not_done:
            for ds in self.task_datastores:
                self.pathspec_index_cache[ds.pathspec_index] = ds
                self.pathspec_cache[ds.pathspec] = ds

    def get_with_pathspec(self, pathspec):
        return self.pathspec_cache.get(pathspec, None)

    def get_with_pathspec_index(self, pathspec_index):
        return self.pathspec_index_cache.get(pathspec_index, None)

    def __iter__(self):
        for ds in self.task_datastores:
            yield ds
