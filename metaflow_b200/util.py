"""Small helpers the datastore path uses (metaflow/util.py:111-147, Path wrapper :45)."""
from io import BytesIO


class Path(object):
    """Marks a string as a filesystem path (as opposed to log content) for save_logs."""

    def __init__(self, path):
        self.path = path

    def __str__(self):
        return self.path


def is_stringish(x):
    return isinstance(x, (bytes, str))


def to_bytes(x):
    if isinstance(x, str):
        return x.encode("utf-8")
    if isinstance(x, bytes):
        return x
    if isinstance(x, float):
        return repr(x).encode("utf-8")
    return str(x).encode("utf-8")


def to_fileobj(x):
    return BytesIO(to_bytes(x))
