"""Minimal Parameter marker (metaflow/parameters.py): class-level flow parameters are not artifacts
of a task and are skipped by TaskDataStore.persist (task_datastore.py:776-783)."""


class Parameter(object):
    def __init__(self, name, default=None, help=None, **kwargs):
        self.name = name
        self.default = default
        self.help = help
        self.kwargs = kwargs
