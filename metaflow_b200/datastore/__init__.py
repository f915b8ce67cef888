from .inputs import Inputs
from .flow_datastore import FlowDataStore
from .datastore_set import TaskDataStoreSet
from .task_datastore import TaskDataStore
from .spin_datastore import SpinTaskDatastore
from .content_addressed_store import ContentAddressedStore, BlobCache
from .exceptions import DataException, UnpicklableArtifactException
