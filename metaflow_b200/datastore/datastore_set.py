"""Batch-open many task datastores and prefetch artifacts shared by a join into one blob cache.

metaflow/datastore/datastore_set.py:15-85.  The prefetch is the reference's only batched load
(used for join fan-in, task.py:252-279); here it is exactly one GpuCodec.unpack_batch over all
distinct keys, which is what turns a join's N serial gunzips into one GPU batch.
"""
from .content_addressed_store import BlobCache


class ImmutableBlobCache(BlobCache):
    """Serves only what was preloaded; never stores (datastore_set.py:76-85)."""

    def __init__(self, preloaded):
        self._blobs = dict(preloaded)

    def load_key(self, key):
        return self._blobs.get(key)

    def store_key(self, key, blob):
        return None


class TaskDataStoreSet(object):
    def __init__(self, flow_datastore, run_id, steps=None, pathspecs=None, prefetch_data_artifacts=None,
                 allow_not_done=False, join_type=None, orig_flow_datastore=None, spin_artifacts=None):
        self.task_datastores = flow_datastore.get_task_datastores(
            run_id, steps=steps, pathspecs=pathspecs, allow_not_done=allow_not_done, join_type=join_type,
            orig_flow_datastore=orig_flow_datastore, spin_artifacts=spin_artifacts)
        if prefetch_data_artifacts:
            self._prefetch(flow_datastore.ca_store, prefetch_data_artifacts)
        done = [] if allow_not_done else self.task_datastores       # unfinished tasks have no stable pathspec index
        self.pathspec_index_cache = {ds.pathspec_index: ds for ds in done}
        self.pathspec_cache = {ds.pathspec: ds for ds in done}

    def _prefetch(self, ca_store, names):
        """All distinct keys behind `names` in all tasks -> ONE batched GPU unpack -> a read-only blob cache."""
        keys = {k for ds in self.task_datastores for k in ds.keys_for_artifacts(names) if k is not None}
        ca_store.set_blob_cache(ImmutableBlobCache(ca_store.load_blobs(keys)))

    def get_with_pathspec(self, pathspec):
        return self.pathspec_cache.get(pathspec)

    def get_with_pathspec_index(self, pathspec_index):
        return self.pathspec_index_cache.get(pathspec_index)

    def __iter__(self):
        return iter(self.task_datastores)
