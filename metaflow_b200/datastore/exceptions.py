"""metaflow/datastore/exceptions.py:4-13 shapes, re-exported from metaflow_b200.exception."""
from ..exception import DataException, UnpicklableArtifactException  # noqa: F401
