"""Local-filesystem storage backend (`--datastore=local`).

On-disk behaviour of metaflow/plugins/datastores/local_storage.py:12-162: blob at
<root>/<path>, JSON sidecar <path>_meta, both written to a temp file in the target directory and
renamed into place (:114-129) so concurrent writers of the same key are harmless; datastore
root discovery walks up from the cwd looking for `.metaflow` (:19-53).
"""
import errno
import json
import os
import tempfile
from concurrent.futures import ThreadPoolExecutor

from ...datastore.datastore_storage import CloseAfterUse, DataStoreStorage
from ...metaflow_config import DATASTORE_LOCAL_DIR, DATASTORE_SYSROOT_LOCAL


class LocalStorage(DataStoreStorage):
    TYPE = "local"
    METADATA_DIR = "_meta"
    DATASTORE_DIR = DATASTORE_LOCAL_DIR
    SYSROOT_VAR = DATASTORE_SYSROOT_LOCAL
    IO_THREADS = 8

    @classmethod
    def get_datastore_root_from_config(cls, echo, create_on_absent=True):
        if cls.SYSROOT_VAR is not None:
            return os.path.join(cls.SYSROOT_VAR, cls.DATASTORE_DIR)
        here = os.getcwd()
        first_choice = os.path.realpath(os.path.join(here, cls.DATASTORE_DIR))
        candidate = first_choice
        while not os.path.isdir(candidate):
            parent = os.path.dirname(here)
            if parent == here:  # hit the filesystem root without finding one
                if not create_on_absent:
                    return None
                echo("Creating %s datastore in current directory (%s)" % (cls.TYPE, first_choice))
                os.mkdir(first_choice)
                return first_choice
            here = parent
            candidate = os.path.join(here, cls.DATASTORE_DIR)
        return candidate

    @staticmethod
    def _makedirs(path):
        try:
            os.makedirs(path)
        except OSError as e:
            if e.errno != errno.EEXIST:
                raise

    @staticmethod
    def _atomic_write(full_path, data, mode="wb"):
        fd, tmp = tempfile.mkstemp(dir=os.path.dirname(full_path))
        try:
            with os.fdopen(fd, mode) as f:
                for piece in getattr(data, "views", (data,)):     # a striped blob is written extent by extent
                    f.write(piece)
            os.rename(tmp, full_path)
            tmp = None
        finally:
            if tmp is not None:
                try:
                    os.unlink(tmp)
                except OSError:
                    pass

    def is_file(self, paths):
        return [os.path.isfile(self.full_uri(p)) for p in paths]

    def info_file(self, path):
        if not self.is_file([path])[0]:
            return False, None
        try:
            with open("%s_meta" % self.full_uri(path), "r") as f:
                return True, json.load(f)
        except OSError:
            return True, None

    def size_file(self, path):
        if not self.is_file([path])[0]:
            return None
        try:
            return os.path.getsize(self.full_uri(path))
        except OSError:
            return None

    def list_content(self, paths):
        out = []
        for path in paths:
            if path == self.METADATA_DIR:
                continue
            try:
                names = os.listdir(self.full_uri(path))
            except FileNotFoundError:
                continue
            for name in names:
                if name == self.METADATA_DIR:
                    continue
                child = self.path_join(path, name)
                out.append(self.list_content_result(path=child, is_file=self.is_file([child])[0]))
        return out

    def save_bytes(self, path_and_bytes_iter, overwrite=False, len_hint=0):
        for path, obj in path_and_bytes_iter:
            byte_obj, metadata = obj if isinstance(obj, tuple) else (obj, None)
            full = self.full_uri(path)
            if not overwrite and os.path.exists(full):
                continue
            self._makedirs(os.path.dirname(full))
            self._atomic_write(full, byte_obj.read())
            if metadata:
                self._atomic_write("%s_meta" % full, json.dumps(metadata).encode("utf-8"))

    def save_views(self, path_view_meta_iter):
        """Batched-writer extension used by the GPU content-addressed store: items are
        (path, buffer, metadata) where `buffer` is any bytes-like (e.g. a view of the pinned output
        arena), written with the same temp-file + rename protocol without an intermediate copy.
        A batch is written by a small thread pool: write(2) releases the GIL, and at GPU speeds the
        per-blob filesystem cost is what is left (hard part 5)."""
        items = list(path_view_meta_iter)

        def write_one(item):
            path, view, metadata = item
            full = self.full_uri(path)
            self._makedirs(os.path.dirname(full))
            self._atomic_write(full, view)
            if metadata:
                self._atomic_write("%s_meta" % full, json.dumps(metadata).encode("utf-8"))

        if len(items) < 4 or sum(len(v) for _, v, _ in items) < (8 << 20):
            for it in items:
                write_one(it)
            return
        with ThreadPoolExecutor(max_workers=self.IO_THREADS) as pool:
            list(pool.map(write_one, items))

    def read_files(self, full_paths):
        """[local path] -> [bytes], read by the same thread pool (used by the batched loader)."""
        def read_one(p):
            with open(p, "rb") as f:
                return f.read()

        if len(full_paths) < 4:
            return [read_one(p) for p in full_paths]
        with ThreadPoolExecutor(max_workers=self.IO_THREADS) as pool:
            return list(pool.map(read_one, full_paths))

    def load_bytes(self, paths):
        def results():
            for path in paths:
                full = self.full_uri(path)
                if not os.path.exists(full):
                    yield path, None, None
                    continue
                meta = None
                if os.path.exists("%s_meta" % full):
                    with open("%s_meta" % full, "r") as f:
                        meta = json.load(f)
                yield path, full, meta

        return CloseAfterUse(results())
