"""Read-through wrapper used by `spin` (metaflow/datastore/spin_datastore.py:1-91): artifacts the
user supplied override those of the original task's datastore; everything else is delegated."""
from itertools import chain


class SpinTaskDatastore(object):
    def __init__(self, flow_name, run_id, step_name, task_id, orig_datastore, spin_artifacts):
        self.flow_name, self.run_id, self.step_name, self.task_id = flow_name, run_id, step_name, task_id
        self.orig_datastore = orig_datastore
        self.spin_artifacts = spin_artifacts or {}
        self._task = None
        self._objects = dict(orig_datastore._objects)
        self._info = dict(orig_datastore._info)
        for name in self.spin_artifacts:
            self._objects.setdefault(name, None)

    def __getitem__(self, name):
        if name in self.spin_artifacts:
            return self.spin_artifacts[name]
        return self.orig_datastore[name]

    def __contains__(self, name):
        return name in self.spin_artifacts or name in self.orig_datastore

    def get(self, name, default=None):
        return self[name] if name in self else default

    def items(self):
        return chain(((k, None) for k in self.spin_artifacts), self.orig_datastore.items())

    def __getattr__(self, name):
        return getattr(self.orig_datastore, name)
