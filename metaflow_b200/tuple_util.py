"""ForeachFrame: one level of a foreach nesting (metaflow/tuple_util.py:19-27); stored in the
`_foreach_stack` artifact and read back by TaskDataStore.pathspec_index."""
from collections import namedtuple

ForeachFrame = namedtuple("ForeachFrame", ["step", "var", "num_splits", "index", "value"])
