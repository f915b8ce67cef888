"""Exception shapes the datastore path raises (mirrors metaflow/exception.py:43-60,107-108
and metaflow/datastore/exceptions.py:4-13 so callers' `except` clauses keep working)."""


class MetaflowException(Exception):
    headline = "Flow failed"

    def __init__(self, msg="", lineno=None, source_file=None):
        self.message = msg
        self.line_no = lineno
        self.source_file = source_file
        super().__init__()

    def __str__(self):
        where = ""
        if self.source_file:
            where = "%s:" % self.source_file
        if self.line_no:
            where = "line %d:" % self.line_no
        return ("%s: " % where if where else "") + str(self.message)


class MetaflowInternalError(MetaflowException):
    headline = "Internal error"


class DataException(MetaflowException):
    headline = "Data store error"


class UnpicklableArtifactException(MetaflowException):
    headline = "Cannot pickle artifact"

    def __init__(self, artifact_name):
        super().__init__(msg='Cannot pickle dump artifact named "%s"' % artifact_name, lineno=None)


class GpuCasUnavailable(MetaflowException):
    """Raised (loudly) when the sm_100a library or a CUDA device is missing.
    The product has no CPU fallback for the hot path."""

    headline = "GPU content-addressed store unavailable"
