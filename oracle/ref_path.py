"""CPU restatement of the reference's artifact hot path (TEST INFRASTRUCTURE ONLY).

The reference (Netflix/metaflow 2.19.24) is pure Python; the arithmetic of the
path lives in third-party natives that are NOT under /root/reference:

  * SHA-1           -> hashlib.sha1  (OpenSSL 3.0.13 in this image), FIPS 180-4
  * gzip / DEFLATE  -> gzip.GzipFile -> zlib 1.3, RFC 1952 / RFC 1951, level 3
  * CRC-32 / ISIZE  -> zlib.crc32 inside gzip.GzipFile, RFC 1952 section 8
  * pickle          -> CPython _pickle, protocol 4

The reference pins none of these versions (setup.py install_requires is
["requests", "boto3"]); they are whatever the interpreter ships.  This module
restates the reference's *call sequence* on those same natives, function by
function, each citing the reference file:line it follows.  The independent
restatement of the published algorithms themselves (SHA-1, CRC-32, INFLATE,
RFC 1952 framing) is in oracle/mfcas_oracle.c.

Pinning: tests/golden/ holds blobs + keys written by the real reference
imported from /root/reference (oracle/make_golden.py); tests/test_oracle.py
checks this module and the C oracle against them.  The reference's own tests
hold no golden digests/bytes for this path, so those
generated fixtures plus FIPS/RFC known answers are the pin.
"""
import gzip
import json
import os
import pickle
import tempfile
import zlib
from hashlib import sha1
from io import BytesIO

CAS_META_V1 = {"cas_raw": False, "cas_version": 1}


def pickle_artifact(obj):
    """task_datastore.py:350-378 -- pickle.dumps(obj, protocol=4) and the _info record."""
    blob = pickle.dumps(obj, protocol=4)
    info = {"size": len(blob), "type": str(type(obj)), "encoding": "gzip+pickle-v4"}
    return blob, info


def blob_key(blob):
    """content_addressed_store.py:98 -- sha1(blob).hexdigest()."""
    return sha1(blob).hexdigest()


def cas_path(prefix, key):
    """content_addressed_store.py:99 -- path_join(prefix, key[:2], key)."""
    return "/".join((prefix.rstrip("/"), key[:2], key))


def pack_v1(blob, mtime=None):
    """content_addressed_store.py:211-216 -- gzip.GzipFile(compresslevel=3) over a BytesIO."""
    buf = BytesIO()
    with gzip.GzipFile(fileobj=buf, mode="wb", compresslevel=3, mtime=mtime) as f:
        f.write(blob)
    return buf.getvalue()


def unpack_v1(packed):
    """content_addressed_store.py:218-220 -- gzip.GzipFile(mode='rb').read() (multi-member, CRC+ISIZE verified)."""
    with gzip.GzipFile(fileobj=BytesIO(packed), mode="rb") as f:
        return f.read()


def crc32(blob):
    """gzip.py:302-314 (_write_raw): self.crc = zlib.crc32(data, self.crc)."""
    return zlib.crc32(blob) & 0xFFFFFFFF


def _atomic_write(full_path, data):
    """plugins/datastores/local_storage.py:114-129 -- mkstemp in the target dir + rename."""
    fd, tmp = tempfile.mkstemp(dir=os.path.dirname(full_path))
    with os.fdopen(fd, "wb") as f:
        f.write(data)
    os.rename(tmp, full_path)


def save_blobs(root, prefix, blobs, raw=False):
    """content_addressed_store.py:41-120 + local_storage.py:131-147.

    Returns [(uri_or_None, key)] in input order; writes <root>/<prefix>/<k[:2]>/<k>
    and the <path>_meta sidecar for blobs whose key is not present yet.
    """
    out = []
    for blob in blobs:
        key = blob_key(blob)
        path = cas_path(prefix, key)
        full = os.path.join(root, path)
        out.append((full if raw else None, key))
        if os.path.isfile(full):  # dedup check, content_addressed_store.py:107
            continue
        os.makedirs(os.path.dirname(full), exist_ok=True)
        _atomic_write(full, blob if raw else pack_v1(blob))
        meta = {"cas_raw": raw, "cas_version": 1}
        _atomic_write(full + "_meta", json.dumps(meta).encode("utf-8"))
    return out


def load_blobs(root, prefix, keys, force_raw=False):
    """content_addressed_store.py:122-204 + local_storage.py:149-162 -> [(key, bytes)]."""
    out = []
    for key in keys:
        full = os.path.join(root, cas_path(prefix, key))
        meta = None
        if os.path.exists(full + "_meta"):
            with open(full + "_meta") as f:
                meta = json.load(f)
        with open(full, "rb") as f:
            data = f.read()
        if force_raw or (meta and meta.get("cas_raw", False)):
            out.append((key, data))
        else:
            out.append((key, unpack_v1(data)))
    return out


def save_artifacts(root, flow_name, named_objects):
    """task_datastore.py:329-383 -> (objects{name:key}, info{name:{...}})."""
    objects, info = {}, {}
    names, blobs = [], []
    for name, obj in named_objects:
        blob, inf = pickle_artifact(obj)
        info[name] = inf
        names.append(name)
        blobs.append(blob)
    for name, (_, key) in zip(names, save_blobs(root, flow_name + "/data", blobs)):
        objects[name] = key
    return objects, info


def load_artifacts(root, flow_name, objects, names):
    """task_datastore.py:386-442 -> [(name, obj)], one fresh unpickle per name."""
    by_key = {}
    for n in names:
        by_key.setdefault(objects[n], []).append(n)
    out = []
    for key, blob in load_blobs(root, flow_name + "/data", list(by_key)):
        for n in by_key[key]:
            out.append((n, pickle.loads(blob)))
    return out
