"""Storage-backend interface below the content-addressed store.

Same contract as metaflow/datastore/datastore_storage.py:7-273 (CloseAfterUse :7-23,
DataStoreStorage :26, path helpers :106-125, is_file :127, info_file :144, size_file :162,
list_content :178, save_bytes :206, load_bytes :243).  This seam sits BELOW pack/hash: it only
ever sees bytes that are already gzip-packed, so stock Metaflow backends plug in unchanged.
"""
import re
from collections import namedtuple

from ..exception import DataException


class CloseAfterUse(object):
    """`with CloseAfterUse(data, closer) as data:` -- closer.close() runs when the block exits."""

    def __init__(self, data, closer=None):
        self.data = data
        self._closer = closer

    def __enter__(self):
        return self.data

    def __exit__(self, *exc):
        if self._closer is not None:
            self._closer.close()


class DataStoreStorage(object):
    TYPE = None
    datastore_root = None
    path_rexp = None

    list_content_result = namedtuple("list_content_result", "path is_file")

    def __init__(self, root=None):
        self.datastore_root = root if root else self.datastore_root

    @classmethod
    def get_datastore_root_from_config(cls, echo, create_on_absent=True):
        raise NotImplementedError

    @classmethod
    def get_datastore_root_from_location(cls, path, flow_name):
        """Recover the datastore root from a CAS blob location (datastore_storage.py:61-104)."""
        if cls.path_rexp is None:
            cls.path_rexp = re.compile(
                cls.path_join(
                    "(?P<root>.*)",
                    "(?P<flow_name>[_a-zA-Z][_a-zA-Z0-9]+)",
                    "data",
                    "(?P<init>[0-9a-f]{2})",
                    "(?:r_)?(?P=init)[0-9a-f]{38}",
                )
            )
        m = cls.path_rexp.match(path)
        if m is None or m.group("flow_name") != flow_name:
            raise DataException(
                "Location '%s' does not correspond to a valid location for flow '%s'." % (path, flow_name)
            )
        return m.group("root")

    @classmethod
    def path_join(cls, *components):
        if not components:
            return ""
        parts = [components[0].rstrip("/")]
        parts.extend(c.strip("/") for c in components[1:])
        return "/".join(parts)

    @classmethod
    def path_split(cls, path):
        return path.split("/")

    @classmethod
    def basename(cls, path):
        return path.split("/")[-1]

    @classmethod
    def dirname(cls, path):
        return path.rsplit("/", 1)[0]

    def full_uri(self, path):
        return self.path_join(self.datastore_root, path)

    # -- to be provided by a backend -------------------------------------------------
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
