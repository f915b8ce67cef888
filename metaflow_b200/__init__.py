"""metaflow_b200: a B200-native implementation of ONE hot path of Netflix/metaflow -- the artifact
datastore (TaskDataStore.save_artifacts/load_artifacts -> ContentAddressedStore.save_blobs/load_blobs:
pickle -> SHA-1 key -> gzip -> local CAS and back) -- behind the reference's own datastore surface.
See DESIGN.md for scope and INTEGRATION.md for the C-ABI boundary."""
from .exception import (DataException, GpuCasUnavailable, MetaflowException, MetaflowInternalError,
                        UnpicklableArtifactException)
from .datastore import (ContentAddressedStore, FlowDataStore, Inputs, TaskDataStore, TaskDataStoreSet)
from .plugins.datastores.local_storage import LocalStorage

from .flowspec import FlowRunner, FlowSpec, step

__version__ = "0.1.0"


def __getattr__(name):            # `metaflow_b200.FileCache` without importing the client package up front
    if name in ("FileCache", "FileBlobCache", "TaskMetadataCache"):
        from . import client

        return getattr(client, name)
    raise AttributeError("module 'metaflow_b200' has no attribute '%s'" % name)
