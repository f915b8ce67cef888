"""Test-only codec: the CPU oracle behind the GpuCodec interface, so the HOST logic of the
datastore classes (batch windows, dedup, file layout, error text) can be tested without a GPU.
Never imported by the product."""
from oracle import ref_path


class OracleCodec(object):
    def __init__(self):
        self.calls = {"pack": 0, "hash": 0, "unpack": 0}
        self.packed_blobs = 0

    def hash_batch(self, blobs):
        self.calls["hash"] += 1
        return [ref_path.blob_key(bytes(b)) for b in blobs]

    def pack_batch(self, blobs):
        self.calls["pack"] += 1
        self.packed_blobs += len(blobs)
        return [ref_path.blob_key(bytes(b)) for b in blobs], [memoryview(ref_path.pack_v1(bytes(b), mtime=0)) for b in blobs]

    def unpack_batch(self, packed, as_bytes=True):
        self.calls["unpack"] += 1
        out = []
        for p in packed:
            try:
                out.append(ref_path.unpack_v1(bytes(p)))
            except Exception as e:  # noqa: BLE001 -- mirror GpuCodec: errors are returned per blob
                out.append(e)
        return out
