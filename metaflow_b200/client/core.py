"""Client object model over the file cache: Flow / Run / Step / Task / DataArtifact / MetaflowData.

The read side of the hot path as a user of the Client API sees it (metaflow/client/core.py:973-1010
`DataArtifact.data`, :1106-1149 `MetaflowData`, :1511 `Task.data`, :2257 `Run.data`): every artifact access ends in
`FileCache.get_artifact*` -> `TaskDataStore.load_artifacts` -> `ContentAddressedStore.load_blobs` -> GPU unpack, with
the on-disk client cache in between.  The reference resolves objects through a metadata provider; that service is
out of scope here, so objects are discovered from the datastore layout itself (`<root>/<Flow>/<run>/<step>/<task>/
<attempt>.data.json`), which is what the local metadata provider mirrors anyway.

Added for the GPU path: loads are BATCHED wherever the object model knows more than one artifact is wanted --
`Task.data` preloads nothing until touched but `MetaflowData.preload()`, `Step.artifacts(names)` and `Run.data`
over a foreach step fetch all tasks' artifacts with one `FileCache.get_artifacts_for_tasks` (one GPU unpack).
"""
import json
import os

from ..exception import MetaflowException
from ..plugins.datastores.local_storage import LocalStorage
from .filecache import FileCache

filecache = None
_datastore_root = None


class MetaflowNotFound(MetaflowException):
    headline = "Object not found"


def set_datastore_root(root):
    """Directory that holds the `<Flow>/...` trees (what `--datastore-root` names; `.metaflow` by default)."""
    global _datastore_root
    _datastore_root = root


def get_datastore_root():
    if _datastore_root is not None:
        return _datastore_root
    return LocalStorage.get_datastore_root_from_config(lambda *_: None, create_on_absent=False)


def set_filecache(fc):
    global filecache
    filecache = fc


def _filecache():
    global filecache
    if filecache is None:
        filecache = FileCache()
    return filecache


def _listdir(path):
    try:
        return sorted(os.listdir(path))
    except OSError:
        return []


def _natural(name):
    return (0, int(name)) if name.isdigit() else (1, name)


class MetaflowObject(object):
    _NAME, _DEPTH = "base", 0

    def __init__(self, pathspec, _root=None):
        self._root = _root if _root is not None else get_datastore_root()
        if self._root is None:
            raise MetaflowNotFound("No local datastore root found: call client.set_datastore_root()")
        self._components = pathspec.split("/")
        if len(self._components) != self._DEPTH:
            raise MetaflowException("%s needs a pathspec of %d components, got '%s'" % (type(self).__name__, self._DEPTH, pathspec))
        self.pathspec = pathspec
        self.id = self._components[-1]
        if not self._exists():
            raise MetaflowNotFound("%s('%s') does not exist" % (type(self).__name__, pathspec))

    @property
    def path_components(self):
        return list(self._components)

    def _dir(self):
        return os.path.join(self._root, *self._components)

    def _exists(self):
        return os.path.isdir(self._dir())

    def _child_ids(self):
        skip = {LocalStorage.METADATA_DIR, "data"} if self._DEPTH == 1 else {LocalStorage.METADATA_DIR}
        return sorted((n for n in _listdir(self._dir()) if n not in skip and os.path.isdir(os.path.join(self._dir(), n))),
                      key=_natural)

    def _child(self, child_id):
        return self._CHILD("%s/%s" % (self.pathspec, child_id), _root=self._root)

    def __iter__(self):
        for cid in reversed(self._child_ids()):                 # newest first, as the reference client iterates
            try:
                yield self._child(cid)
            except MetaflowNotFound:
                continue

    def __getitem__(self, child_id):
        if child_id not in self._child_ids():
            raise KeyError(child_id)
        return self._child(str(child_id))

    def __contains__(self, child_id):
        return str(child_id) in self._child_ids()

    def __repr__(self):
        return "%s('%s')" % (type(self).__name__, self.pathspec)


class DataArtifact(MetaflowObject):
    """Leaf of the hierarchy: one named artifact of one task attempt."""

    _NAME, _DEPTH = "artifact", 5

    def __init__(self, pathspec, _root=None, _task=None):
        self._task = _task
        super().__init__(pathspec, _root=_root)

    def _exists(self):
        if self._task is None:
            self._task = Task("/".join(self._components[:4]), _root=self._root)
        return self.id in self._task._meta()["objects"]

    @property
    def sha(self):
        return self._task._meta()["objects"][self.id]

    @property
    def size(self):
        return self._task._meta()["info"].get(self.id, {}).get("size", 0)

    @property
    def data(self):
        """The unpickled object (client/core.py:973-1010): through the file cache, then the GPU datastore."""
        flow, run, step, task = self._components[:4]
        meta = self._task._meta()
        one = {"objects": {self.id: meta["objects"][self.id]}, "info": {self.id: meta["info"].get(self.id, {"encoding": "gzip+pickle-v4"})}}
        return _filecache().get_artifact(LocalStorage.TYPE, self._root, one, flow, run, step, task, self.id)


class MetaflowData(object):
    """`Task.data`: artifacts by dot notation, loaded lazily; `preload()` fetches many with one batched load."""

    def __init__(self, task, names):
        self.__dict__["_task"] = task
        self.__dict__["_names"] = list(names)
        self.__dict__["_loaded"] = {}

    def preload(self, names=None):
        names = [n for n in (names or self._names) if n in self._names and n not in self._loaded]
        if names:
            flow, run, step, task = self._task.path_components
            got = _filecache().get_artifacts_for_tasks(LocalStorage.TYPE, self._task._root, flow,
                                                       [(run, step, task, self._task._meta())], names)
            self._loaded.update(got[(run, step, task)])
        return self

    def __getattr__(self, name):
        if name not in self._names:
            raise AttributeError(name)
        if name not in self._loaded:
            self._loaded[name] = self._task[name].data
        return self._loaded[name]

    def __contains__(self, name):
        return name in self._names

    def __repr__(self):
        return "<MetaflowData: %s>" % ", ".join(self._names)


class Task(MetaflowObject):
    _NAME, _DEPTH = "task", 4

    def __init__(self, pathspec, attempt=None, _root=None):
        self._attempt = attempt
        self._meta_cache = None
        super().__init__(pathspec, _root=_root)

    def _meta(self):
        """{'objects': name->key, 'info': name->info} of the latest DONE attempt (or the requested one)."""
        if self._meta_cache is None:
            files = _listdir(self._dir())
            attempts = sorted(int(f.split(".", 1)[0]) for f in files if f.endswith(".DONE.lock"))
            if self._attempt is not None:
                attempts = [a for a in attempts if a == self._attempt]
            if not attempts:
                raise MetaflowNotFound("No completed attempt of task '%s'" % self.pathspec)
            self._attempt = attempts[-1]
            with open(os.path.join(self._dir(), "%d.data.json" % self._attempt)) as f:
                dj = json.load(f)
            self._meta_cache = {"objects": dj.get("objects", {}), "info": dj.get("info", {})}
        return self._meta_cache

    @property
    def current_attempt(self):
        self._meta()
        return self._attempt

    def _user_names(self):
        return [n for n in self._meta()["objects"] if not n.startswith("_")]

    @property
    def data(self):
        return MetaflowData(self, self._user_names())

    @property
    def artifacts(self):
        return [DataArtifact("%s/%s" % (self.pathspec, n), _root=self._root, _task=self) for n in self._user_names()]

    def __iter__(self):
        return iter(self.artifacts)

    def __getitem__(self, name):
        if name not in self._meta()["objects"]:
            raise KeyError(name)
        return DataArtifact("%s/%s" % (self.pathspec, name), _root=self._root, _task=self)

    def __contains__(self, name):
        return name in self._meta()["objects"]

    @property
    def finished(self):
        try:
            self._meta()
            return True
        except MetaflowNotFound:
            return False

    @property
    def successful(self):
        """`_task_ok` artifact, as the reference decides it (client/core.py Task.successful)."""
        try:
            return bool(self["_task_ok"].data)
        except (KeyError, MetaflowNotFound):
            return False

    @property
    def index(self):
        try:
            stack = self["_foreach_stack"].data
        except KeyError:
            return None
        return stack[-1].index if stack else None


class Step(MetaflowObject):
    _NAME, _DEPTH, _CHILD = "step", 3, Task

    @property
    def task(self):
        for t in self:
            return t
        return None

    def tasks(self):
        return list(self)

    def artifacts(self, names):
        """{task_id: {name: object}} for ALL tasks of the step with ONE batched load (one GPU unpack): what a
        notebook does after a wide foreach (`[t.data.x for t in step]`) without one decode call per task."""
        flow, run, step = self._components
        tasks = [t for t in self if t.finished]
        specs = [(run, step, t.id, t._meta()) for t in tasks]
        got = _filecache().get_artifacts_for_tasks(LocalStorage.TYPE, self._root, flow, specs, list(names))
        return {t.id: got[(run, step, t.id)] for t in tasks}


class Run(MetaflowObject):
    _NAME, _DEPTH, _CHILD = "run", 2, Step

    def _child_ids(self):
        return [n for n in super()._child_ids() if n != "_parameters"]

    def steps(self):
        return list(self)

    @property
    def end_task(self):
        try:
            return self["end"].task
        except KeyError:
            return None

    @property
    def finished(self):
        t = self.end_task
        return bool(t and t.finished)

    @property
    def successful(self):
        t = self.end_task
        return bool(t and t.successful)

    @property
    def data(self):
        """Artifacts of the `end` task (client/core.py:2257), None while the run has not finished."""
        t = self.end_task
        return t.data if t is not None and t.finished else None


class Flow(MetaflowObject):
    _NAME, _DEPTH, _CHILD = "flow", 1, Run

    def runs(self):
        return list(self)

    @property
    def latest_run(self):
        for r in self:
            return r
        return None

    @property
    def latest_successful_run(self):
        for r in self:
            if r.successful:
                return r
        return None


Task._CHILD = DataArtifact
