"""Tiny driver for profiling: packs 32 x 2 MiB blobs a few times (SHA-1 kernel dominated)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metaflow_b200.codec import GpuCodec
c = GpuCodec()
rng = np.random.default_rng(0)
blobs = [rng.standard_normal(1 << 19, dtype=np.float32).tobytes() for _ in range(32)]
for _ in range(3):
    keys, packed = c.pack_batch(blobs)
import hashlib
assert keys == [hashlib.sha1(b).hexdigest() for b in blobs]
print("ok")
