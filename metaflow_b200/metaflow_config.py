"""The handful of configuration knobs the datastore path reads.

Same precedence idea as metaflow/metaflow_config_funcs.py:84-140 (`from_conf`): the
environment variable METAFLOW_<NAME> wins over the default.  (Profile JSON files are
part of the reference's config subsystem, which is out of scope here.)
"""
import json
import os


def from_conf(name, default=None, validate_fn=None):
    raw = os.environ.get("METAFLOW_%s" % name)
    if raw is None:
        return default
    if default is not None and not isinstance(default, str):
        try:
            raw = type(default)(json.loads(raw)) if not isinstance(default, bool) else raw.lower() in ("1", "true", "yes")
        except (ValueError, TypeError):
            raise ValueError("METAFLOW_%s has an invalid value %r" % (name, raw))
    if validate_fn:
        validate_fn(name, raw)
    return raw


# metaflow_config.py:29,39,98,604,525
DATASTORE_LOCAL_DIR = ".metaflow"
DEFAULT_DATASTORE = from_conf("DEFAULT_DATASTORE", "local")
DATASTORE_SYSROOT_LOCAL = from_conf("DATASTORE_SYSROOT_LOCAL")
MAX_ATTEMPTS = from_conf("MAX_ATTEMPTS", 6)
PROFILE_FROM_START = from_conf("PROFILE_FROM_START", False)

# New knobs for the GPU path (SURVEY.md section 5 proposal).
GPU_CAS_DEVICES = from_conf("GPU_CAS_DEVICES", "")            # "" = all visible; else "0,1,.."
GPU_CAS_WAVE_BYTES = from_conf("GPU_CAS_WAVE_BYTES", 1 << 30)  # bytes (in + out) staged per pipelined GPU wave
# bytes of blobs gathered from a save/load iterator before one codec call is issued.  SHA-1 is a serial chain per
# blob (~8 ms per MiB on a GPU lane), so a call is only as fast as it is WIDE: the window should hold as many
# blobs as the caller has (the page-locked staging arenas grow on demand up to in + out of this size).
GPU_CAS_WINDOW_BYTES = from_conf("GPU_CAS_WINDOW_BYTES", 4 << 30)
# load windows are also bounded by their DECODED size and blob count (highly compressible artifacts inflate 1000x)
GPU_CAS_WINDOW_UNPACKED_BYTES = from_conf("GPU_CAS_WINDOW_UNPACKED_BYTES", 8 << 30)
GPU_CAS_WINDOW_BLOBS = from_conf("GPU_CAS_WINDOW_BLOBS", 1 << 17)
# page-locked staging arenas larger than this are released after the window that needed them
GPU_CAS_ARENA_KEEP = from_conf("GPU_CAS_ARENA_KEEP", 12 << 30)
GPU_CAS_LZ = from_conf("GPU_CAS_LZ", 1)                        # 0 = Huffman-only encoder

# client-side file cache (metaflow/metaflow_config.py:113-121)
CLIENT_CACHE_PATH = from_conf("CLIENT_CACHE_PATH", "/tmp/metaflow_client")
CLIENT_CACHE_MAX_SIZE = from_conf("CLIENT_CACHE_MAX_SIZE", 10000)                         # megabytes
CLIENT_CACHE_MAX_FLOWDATASTORE_COUNT = from_conf("CLIENT_CACHE_MAX_FLOWDATASTORE_COUNT", 50)
