"""Read-through wrapper used by `spin` (metaflow/datastore/spin_datastore.py:1-91): artifacts the
user supplied override those of the original task's datastore; everything else comes from the task that
is being re-run.  `_objects` / `_info` are what TaskDataStore.persist copies into the new task
(task_datastore.py:760-762): the original task's maps minus the control artifact `_transition` -- user
supplied spin artifacts are NOT in them (they are re-saved by persist like any other attribute)."""


class SpinTaskDatastore(object):
    def __init__(self, flow_name, run_id, step_name, task_id, orig_datastore, spin_artifacts):
        self.flow_name, self.run_id, self.step_name, self.task_id = flow_name, run_id, step_name, task_id
        self.orig_datastore = orig_datastore
        self.spin_artifacts = spin_artifacts or {}
        self._task = None
        self._objects = dict(orig_datastore._objects)
        self._info = dict(orig_datastore._info)
        for control in ("_transition",):
            self._objects.pop(control, None)
            self._info.pop(control, None)

    def __getitem__(self, name):
        if name in self.spin_artifacts:
            return self.spin_artifacts[name]
        try:
            return self.orig_datastore[name]
        except (KeyError, AttributeError) as e:
            raise KeyError("Attribute '%s' not found in the previous execution of the tasks for `%s`."
                           % (name, self.step_name)) from e

    def is_none(self, name):
        return self[name] is None

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def get(self, name, default=None):
        try:
            return self[name]
        except KeyError:
            return default

    def items(self):
        return self._objects.items() if self._objects else {}
