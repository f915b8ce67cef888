"""Storage-backend interface below the content-addressed store.

The contract backends implement is the reference's (metaflow/datastore/datastore_storage.py: CloseAfterUse :7-23,
DataStoreStorage :26, path helpers :106-125, is_file :127, info_file :144, size_file :162, list_content :178,
save_bytes :206, load_bytes :243), so stock Metaflow backends plug in unchanged: this seam sits BELOW pack/hash and
only ever sees bytes that are already gzip-packed.

Two batch hooks are added for the GPU path, with generic defaults written in terms of that contract:
`save_views` (packed blobs arrive as views of page-locked memory, possibly several extents per blob) and
`read_files` (a window of files is wanted at once).  LocalStorage overrides both with thread pools.
"""
import functools
import posixpath
import re
from collections import namedtuple
from io import BytesIO

from ..exception import DataException


class CloseAfterUse(object):
    """`with CloseAfterUse(data, closer) as data:` -- closer.close() runs when the block exits."""

    __slots__ = ("data", "_closer")

    def __init__(self, data, closer=None):
        self.data, self._closer = data, closer

    def __enter__(self):
        return self.data

    def __exit__(self, *exc):
        closer, self._closer = self._closer, None
        if closer is not None:
            closer.close()


@functools.lru_cache(maxsize=None)
def _cas_location_pattern(sep):
    """<root>/<flow>/data/<k[:2]>/[r_]<k>: where a content-addressed blob lives (datastore_storage.py:61-104)."""
    key = r"(?P<init>[0-9a-f]{2})%s(?:r_)?(?P=init)[0-9a-f]{38}" % re.escape(sep)
    return re.compile(sep.join(["(?P<root>.*)", "(?P<flow_name>[_a-zA-Z][_a-zA-Z0-9]+)", "data", key]))


class DataStoreStorage(object):
    TYPE = None
    datastore_root = None
    path_rexp = None                      # kept for backends that inspect it; the pattern itself is cached per separator

    list_content_result = namedtuple("list_content_result", "path is_file")

    def __init__(self, root=None):
        if root:
            self.datastore_root = root

    # -- where the datastore lives ----------------------------------------------------------
    @classmethod
    def get_datastore_root_from_config(cls, echo, create_on_absent=True):
        raise NotImplementedError("%s does not say where its datastore root is" % cls.__name__)

    @classmethod
    def get_datastore_root_from_location(cls, path, flow_name):
        """The datastore root that a CAS blob location belongs to; the flow name must match."""
        hit = _cas_location_pattern("/").match(path)
        if hit is None or hit.group("flow_name") != flow_name:
            raise DataException(
                "Location '%s' does not correspond to a valid location for flow '%s'." % (path, flow_name)
            )
        return hit.group("root")

    # -- paths are '/'-separated strings relative to the root --------------------------------
    @classmethod
    def path_join(cls, *components):
        if not components:
            return ""
        head, rest = components[0].rstrip("/"), [c.strip("/") for c in components[1:]]
        return "/".join([head] + rest)

    @classmethod
    def path_split(cls, path):
        return path.split("/")

    @classmethod
    def basename(cls, path):
        return posixpath.basename(path)

    @classmethod
    def dirname(cls, path):
        return path.rsplit("/", 1)[0]

    def full_uri(self, path):
        return self.path_join(self.datastore_root, path)

    # -- the contract a backend provides ------------------------------------------------------
    def is_file(self, paths):
        """[path] -> [bool]; directories are not files."""
        raise NotImplementedError

    def info_file(self, path):
        """path -> (exists, metadata dict or None)"""
        raise NotImplementedError

    def size_file(self, path):
        """path -> size in bytes or None"""
        raise NotImplementedError

    def list_content(self, paths):
        """one-level `ls` of each directory -> [list_content_result(path, is_file)]"""
        raise NotImplementedError

    def save_bytes(self, path_and_bytes_iter, overwrite=False, len_hint=0):
        """Iterator of (path, fileobj) or (path, (fileobj, metadata)) -> None."""
        raise NotImplementedError

    def load_bytes(self, keys):
        """[path] -> CloseAfterUse(iterator of (path, local_file_or_None, metadata_or_None))"""
        raise NotImplementedError

    # -- batch hooks of the GPU path (defaults in terms of the contract above) ---------------
    def save_views(self, path_view_meta_iter):
        """Items (path, buffer, metadata): `buffer` is bytes-like or has `.views` (extents to concatenate).
        Default: materialise each blob and hand it to save_bytes."""
        def as_files():
            for path, buf, meta in path_view_meta_iter:
                data = b"".join(bytes(v) for v in buf.views) if hasattr(buf, "views") else bytes(buf)
                yield path, (BytesIO(data), meta)

        self.save_bytes(as_files(), overwrite=True)

    def read_files(self, local_paths):
        """[local file path, as yielded by load_bytes] -> [bytes]."""
        out = []
        for p in local_paths:
            with open(p, "rb") as f:
                out.append(f.read())
        return out
