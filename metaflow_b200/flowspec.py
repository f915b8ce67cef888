"""FlowSpec / @step / foreach / join with an in-process runner whose artifact work is BATCHED per step.

User surface of metaflow/flowspec.py (FlowSpec :266, next :909-1153, foreach_stack :654, input/index,
merge_artifacts :738-860, lazy artifact attributes :599-605) and metaflow/decorators.py:1056 (`step`), so
flow files written for Metaflow run unchanged.  What is replaced is the scheduler
(metaflow/runtime.py:1332-1388 `_queue_task_foreach`, :1163-1316 `_queue_task_join`, :1517-1540
`_launch_workers`: one fork+exec'd Python process per task, each hashing and gzipping its own artifacts on
one CPU core): here the tasks of a step run in this process and the step's WHOLE artifact set -- every
sibling task of a foreach -- is pickled and handed to ContentAddressedStore.save_blobs as one batch
(SHA-1 + DEFLATE + CRC-32 on the GPUs, round-robin over devices); a join loads its inputs' artifacts with one
batched unpack.  The datastore on disk is the reference's, file for file.
"""
import inspect
from itertools import count

from .datastore import FlowDataStore, Inputs
from .exception import MetaflowException, UnpicklableArtifactException
from .serializer import dumps_segments, loads_batch
from .parameters import Parameter
from .plugins.datastores.local_storage import LocalStorage
from .tuple_util import ForeachFrame


class InvalidNextException(MetaflowException):
    headline = "Invalid self.next() transition detected"


class MergeArtifactsException(MetaflowException):
    headline = "Unhandled artifacts in merge"

    def __init__(self, msg, unhandled):
        super().__init__(msg)
        self.artifact_names = unhandled


class MissingInMergeArtifactsException(MetaflowException):
    headline = "Missing artifacts in merge"

    def __init__(self, msg, unhandled):
        super().__init__(msg)
        self.artifact_names = unhandled


def step(f):
    """Marks a method as a step (metaflow/decorators.py:1056-1113)."""
    f.is_step = True
    f.decorators = []
    f.name = f.__name__
    return f


class FlowSpec(object):
    # attributes that never become artifacts (flowspec.py:276-289)
    _EPHEMERAL = {"_EPHEMERAL", "_NON_PARAMETERS", "_datastore", "_cached_input", "_graph", "_flow_state",
                  "_steps", "index", "input", "_join_group"}

    def __init__(self, use_cli=False):
        self.name = self.__class__.__name__
        self._datastore = None
        self._transition = None
        self._cached_input = {}
        self._current_step = None
        self._foreach_stack_frames = []

    # -- lazy artifacts: attribute misses are served from the input datastore (flowspec.py:599-605)
    def __getattr__(self, name):
        ds = self.__dict__.get("_datastore")
        group = self.__dict__.get("_join_group")
        if group is not None and ds is not None and name in ds:
            group.fetch(name)                    # one batched load of `name` for ALL inputs of the join
            if name in self.__dict__:
                return self.__dict__[name]
        if ds is not None and name in ds:
            value = ds[name]
            setattr(self, name, value)
            return value
        raise AttributeError("Flow %s has no attribute '%s'" % (self.__dict__.get("name", "?"), name))

    @property
    def index(self):
        """Index of this task in its innermost foreach (flowspec.py:608)."""
        return self._foreach_stack_frames[-1].index if self._foreach_stack_frames else None

    @property
    def input(self):
        """The item of the foreach list this task works on (flowspec.py:632)."""
        return self._foreach_stack_frames[-1].value if self._foreach_stack_frames else None

    def foreach_stack(self):
        return [(f.index, f.num_splits, f.value) for f in self._foreach_stack_frames]

    def next(self, *dsts, **kwargs):
        """self.next(self.a) | self.next(self.a, self.b) | self.next(self.a, foreach='attr')"""
        foreach = kwargs.pop("foreach", None)
        if kwargs:
            raise InvalidNextException("Step *%s* passes unknown keyword arguments to self.next(): %s"
                                       % (self._current_step, ", ".join(kwargs)))
        if self._transition is not None:
            raise InvalidNextException("Multiple self.next() calls detected in step *%s*" % self._current_step)
        names = []
        for d in dsts:
            name = getattr(d, "__name__", None)
            if name is None or not getattr(getattr(self.__class__, name, None), "is_step", False):
                raise InvalidNextException("In step *%s* an argument of self.next() is not a @step method" % self._current_step)
            names.append(name)
        if not names:
            raise InvalidNextException("Step *%s* calls self.next() without a destination" % self._current_step)
        if foreach is not None:
            if len(names) != 1:
                raise InvalidNextException("Step *%s*: a foreach split goes to exactly one step" % self._current_step)
            try:
                values = list(getattr(self, foreach))
            except Exception:
                raise InvalidNextException("Foreach variable *self.%s* in step *%s* is not iterable" % (foreach, self._current_step))
            if not values:
                raise InvalidNextException("Cannot spawn zero tasks: foreach variable *self.%s* is empty" % foreach)
            self._foreach_var = foreach
            self._foreach_num_splits = len(values)
        self._transition = (names, foreach)

    def merge_artifacts(self, inputs, exclude=None, include=None):
        """Propagate unambiguous artifacts of a join's inputs (flowspec.py:738-860): values are compared by
        their content-address KEYS, never by loading them."""
        include, exclude = include or [], exclude or []
        if include and exclude:
            raise MetaflowException("`exclude` and `include` are mutually exclusive in merge_artifacts")
        available = set()
        for inp in inputs:
            available.update(n for n, _ in inp._datastore.items() if not n.startswith("_") and n != "name")
        missing = [n for n in include if n not in available]
        if missing:
            raise MissingInMergeArtifactsException("Artifacts %s requested in merge_artifacts are missing" % missing, missing)
        wanted = [n for n in sorted(available) if (not include or n in include) and n not in exclude and n not in self.__dict__]
        ambiguous, passdown = [], {}
        for n in wanted:
            holders = [inp for inp in inputs if n in inp._datastore]
            keys = {inp._datastore._objects[n] for inp in holders}
            if len(keys) > 1:
                ambiguous.append(n)
            else:
                passdown[n] = holders[0]
        if ambiguous:
            raise MergeArtifactsException(
                "Step *%s* cannot merge the following artifacts due to them having conflicting values: %s"
                % (self._current_step, ", ".join(ambiguous)), ambiguous)
        self.__dict__.setdefault("_merged_from", {}).update(passdown)


# ----------------------------------------------------------------------------------------- runner
class _Task(object):
    __slots__ = ("step", "task_id", "frames", "splits", "input_ds", "flow", "out_ds")

    def __init__(self, step_name, task_id, frames, splits, input_ds):
        self.step, self.task_id, self.frames, self.splits, self.input_ds = step_name, task_id, frames, splits, input_ds
        self.flow = self.out_ds = None


class _JoinGroup(object):
    """Join fan-in (metaflow/task.py:252-279 + datastore_set.py:38-50): the first time ANY input of a join is asked
    for artifact `name`, the artifact is loaded for ALL inputs with one batched GPU unpack -- a join that reads
    `inp.x for inp in inputs` is one unpack_batch instead of N single-blob loads.  Lazy like the reference (an
    artifact nobody touches is never decoded), bounded (`cap` bytes of pickles per fetch, else inputs load one by
    one) and local to the join: no blob cache is left behind on the shared content-addressed store."""

    def __init__(self, flow_datastore, flows, cap=None):
        from . import metaflow_config

        self.ca_store = flow_datastore.ca_store
        self.flows = flows
        self.cap = int(metaflow_config.GPU_CAS_WINDOW_UNPACKED_BYTES) if cap is None else cap
        self.fetches = 0

    def fetch(self, name):
        holders = {}
        for f in self.flows:
            ds = f.__dict__.get("_datastore")
            if ds is not None and name in ds and name not in f.__dict__:
                holders.setdefault(ds._objects[name], []).append(f)
        if len(holders) < 2 and sum(len(v) for v in holders.values()) < 2:
            return                                   # nothing to batch: the ordinary single load follows
        total = 0
        for key, fs in holders.items():
            total += len(fs) * fs[0]._datastore._info.get(name, {}).get("size", 0)
        if total > self.cap:
            return
        self.fetches += 1
        loader = getattr(self.ca_store, "_load_batches", None)
        if loader is None:
            batches = [list(self.ca_store.load_blobs(list(holders)))]
        else:
            batches = loader(list(holders), False, False, as_bytes=False)
        for batch in batches:
            targets, blobs = [], []
            for key, blob in batch:
                for f in holders[key]:               # every input gets its own freshly unpickled object
                    targets.append(f)
                    blobs.append(blob)
            for f, obj in zip(targets, loads_batch(blobs)):
                setattr(f, name, obj)


class FlowRunner(object):
    """run = FlowRunner(MyFlow, ds_root=...).run(); the datastore it leaves behind is a Metaflow datastore."""

    def __init__(self, flow_cls, ds_root, run_id="1", storage_impl=LocalStorage, codec=None, max_num_splits=100000):
        self.flow_cls = flow_cls
        self.run_id = str(run_id)
        self.flow_datastore = FlowDataStore(flow_cls.__name__, storage_impl=storage_impl, ds_root=ds_root, codec=codec)
        self._task_ids = count(1)
        self.max_num_splits = max_num_splits
        self.finished = {}               # step name -> [task datastores, read mode] in execution order

    def _is_join(self, step_name):
        return len(inspect.signature(getattr(self.flow_cls, step_name)).parameters) == 2    # (self, inputs)

    def _bind(self, task, datastore):
        flow = self.flow_cls(use_cli=False)
        flow._datastore = datastore
        flow._current_step = task.step
        flow._foreach_stack_frames = list(task.frames)
        flow._foreach_stack = list(task.frames)
        return flow

    def _open_read(self, step_name, task_id):
        return self.flow_datastore.get_task_datastore(self.run_id, step_name, str(task_id), attempt=0)

    def _run_step_tasks(self, tasks, join_inputs=None):
        """Run the user code of all sibling tasks of one step and persist ALL their artifacts through ONE
        save_blobs call.  The call pulls a generator: a task's step function runs when the store asks for its
        blobs, its artifacts are pickled by reference (serializer.dumps_segments, the same serialiser
        TaskDataStore.save_artifacts uses) and leave the flow object right away (what persist does,
        task_datastore.py:794-806), so at most one window of artifacts (GPU_CAS_WINDOW_BYTES) is alive at a time
        however wide the foreach is."""
        fds = self.flow_datastore
        staged = []                      # (task, names): which results belong to which task

        def blobs():
            for t in tasks:
                t.out_ds = fds.get_task_datastore(self.run_id, t.step, str(t.task_id), attempt=0, mode="w")
                t.out_ds.init_task()
                t.flow = self._bind(t, t.input_ds)
                fn = getattr(t.flow, t.step)
                if join_inputs is not None:
                    fn(join_inputs)
                else:
                    fn()
                if t.flow._transition is None and t.step != "end":
                    raise InvalidNextException("Step *%s* does not call self.next()" % t.step)
                if t.flow._transition is None:
                    t.flow._transition = ([], None)
                out = t.out_ds
                if t.input_ds is not None:                       # unchanged artifacts pass down by key, no re-save
                    out._objects.update(t.input_ds._objects)
                    out._info.update(t.input_ds._info)
                for n, holder in t.flow.__dict__.pop("_merged_from", {}).items():
                    out.passdown_partial(holder._datastore, [n])
                names = []
                keep = {t.flow._transition[1]}                   # the foreach variable is read again for the fan-out
                for var, val in sorted(t.flow.__dict__.items()):
                    if var.startswith("__") or var in FlowSpec._EPHEMERAL or isinstance(val, Parameter):
                        continue
                    if var == "_foreach_stack_frames":
                        continue
                    try:
                        blob = dumps_segments(val)
                    except TypeError as e:
                        raise UnpicklableArtifactException(var) from e
                    out._info[var] = {"size": len(blob), "type": str(type(val)), "encoding": "gzip+pickle-v4"}
                    names.append(var)
                    yield blob
                for var in names:                                # hand-over: user artifacts leave the flow object
                    if not var.startswith("_") and var != "name" and var not in keep:
                        del t.flow.__dict__[var]
                staged.append((t, names))

        # ---- the batcher: one save_blobs call for the step's whole artifact set (windowed inside the store)
        results = fds.ca_store.save_blobs(blobs(), len_hint=len(tasks))
        pos = 0
        for t, names in staged:
            for n in names:
                t.out_ds._objects[n] = results[pos].key
                pos += 1
            t.out_ds.done()
        return [self._open_read(t.step, t.task_id) for t in tasks]

    def run(self):
        """Breadth-first over steps: all sibling tasks of a step form one batch.  Every task carries the stack
        of splits it is inside of; a join closes the innermost one (a foreach: all siblings of one parent;
        a static split: one task per branch)."""
        frontier = [("start", [_Task("start", next(self._task_ids), [], [], None)], None)]
        arrivals = {}                    # (join step, split id) -> {branch step: (task, datastore)}
        while frontier:
            step_name, tasks, join_inputs = frontier.pop(0)
            done = self._run_step_tasks(tasks, join_inputs)
            self.finished.setdefault(step_name, []).extend(done)
            transitions = {(tuple(t.flow._transition[0]), t.flow._transition[1]) for t in tasks}
            if len(transitions) != 1:
                raise InvalidNextException("Tasks of step *%s* disagree on self.next()" % step_name)
            (next_steps, foreach), = transitions
            if not next_steps:
                continue
            if foreach is not None:                                           # foreach fan-out
                nxt, children = next_steps[0], []
                for t, ds in zip(tasks, done):
                    values = list(t.flow.__dict__[foreach]) if foreach in t.flow.__dict__ else list(ds[foreach])
                    if len(values) > self.max_num_splits:
                        raise MetaflowException("Foreach in step *%s* yields %d splits (max %d)" % (step_name, len(values), self.max_num_splits))
                    for i, v in enumerate(values):
                        frame = ForeachFrame(step_name, foreach, len(values), i, v)
                        children.append(_Task(nxt, next(self._task_ids), t.frames + [frame], t.splits + [("foreach", t.task_id)], ds))
                frontier.append((nxt, children, None))
                continue
            if len(next_steps) > 1:                                           # static split
                for nxt in next_steps:
                    children = [_Task(nxt, next(self._task_ids), t.frames, t.splits + [("static", t.task_id, next_steps)], ds)
                                for t, ds in zip(tasks, done)]
                    frontier.append((nxt, children, None))
                continue
            nxt = next_steps[0]
            if not self._is_join(nxt):                                        # linear
                frontier.append((nxt, [_Task(nxt, next(self._task_ids), t.frames, t.splits, ds) for t, ds in zip(tasks, done)], None))
                continue
            groups = {}                                                       # join: close the innermost split
            for t, ds in zip(tasks, done):
                if not t.splits:
                    raise InvalidNextException("Step *%s* joins but there is no open split" % nxt)
                groups.setdefault(t.splits[-1][:2], []).append((t, ds))
            join_batch = []
            for split_key, members in groups.items():
                if split_key[0] == "foreach":
                    slot = arrivals.setdefault((nxt, split_key), {})
                    for t, ds in members:
                        slot[t.frames[-1].index] = (t, ds)
                    if len(slot) < members[0][0].frames[-1].num_splits:
                        continue                                              # siblings that took another path (inner joins) still to come
                    arrivals.pop((nxt, split_key))
                    members = [slot[i] for i in sorted(slot)]
                    t0 = members[0][0]
                    jt = _Task(nxt, next(self._task_ids), t0.frames[:-1], t0.splits[:-1], None)
                else:
                    branches = members[0][0].splits[-1][2]
                    slot = arrivals.setdefault((nxt, split_key), {})
                    for t, ds in members:
                        slot[t.step] = (t, ds)
                    if len(slot) < len(branches):
                        continue                                              # other branches still running
                    arrivals.pop((nxt, split_key))
                    members = [slot[b] for b in branches if b in slot] + [v for k, v in slot.items() if k not in branches]
                    t0 = members[0][0]
                    jt = _Task(nxt, next(self._task_ids), t0.frames, t0.splits[:-1], None)
                join_batch.append((jt, members))
            # every join task of this step is its own batch entry: its inputs differ
            for jt, members in join_batch:
                flows = [self._bind(t, ds) for t, ds in members]
                group = _JoinGroup(self.flow_datastore, flows)
                for f in flows:
                    f._join_group = group
                frontier.append((nxt, [jt], Inputs(flows)))
        return self

    def task_datastores(self, step_name):
        return self.finished[step_name]
