"""Client read path: the on-disk client cache of
metaflow/client/filecache.py in front of the GPU-backed datastore classes."""
from .filecache import FileBlobCache, FileCache, FileCacheException, TaskMetadataCache
from .core import (DataArtifact, Flow, MetaflowData, MetaflowNotFound, Run, Step, Task, get_datastore_root,
                   set_datastore_root, set_filecache)

__all__ = ["FileCache", "FileBlobCache", "TaskMetadataCache", "FileCacheException", "Flow", "Run", "Step", "Task",
           "DataArtifact", "MetaflowData", "MetaflowNotFound", "set_datastore_root", "get_datastore_root", "set_filecache"]
