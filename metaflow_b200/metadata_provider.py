"""Record shapes passed to an (optional) metadata provider by TaskDataStore.done()
(metaflow/metadata_provider/metadata.py DataArtifact / MetaDatum namedtuples)."""
from collections import namedtuple

DataArtifact = namedtuple("DataArtifact", "name ds_type ds_root url type sha")
MetaDatum = namedtuple("MetaDatum", "field value type tags")
