"""Client read path: the on-disk client cache of
metaflow/client/filecache.py in front of the GPU-backed datastore classes."""
from .filecache import FileBlobCache, FileCache, FileCacheException, TaskMetadataCache

__all__ = ["FileCache", "FileBlobCache", "TaskMetadataCache", "FileCacheException"]
