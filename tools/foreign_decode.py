"""How fast reference-written (stock zlib, no index) artifacts decode: one lane per blob.  Run on a GPU box."""
import gzip, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metaflow_b200.codec import GpuCodec
codec = GpuCodec(devices=[0])
rng = np.random.default_rng(1)
for mib, n in ((1, 256), (16, 64)):
    raw = [rng.standard_normal(mib << 18, dtype=np.float32).tobytes() for _ in range(8)]
    z = [gzip.compress(r, 3, mtime=0) for r in raw]
    blobs = [z[i % 8] for i in range(n)]
    for attempt in range(2):
        t = time.perf_counter(); out = codec.unpack_batch(blobs, as_bytes=False); dt = time.perf_counter() - t
    assert bytes(out[1]) == raw[1]
    print("%d foreign gzip blobs of %d MiB: %.3f s -> %.3f GB/s (%.1f MB/s per blob)" % (n, mib, dt, n * (mib << 20) / dt / 1e9, (mib << 20) / dt / 1e6 * 1))
codec.close()
