"""BASELINE config C3: a wide foreach of small artifacts (65536 x 64 KiB fp32 arrays, pickle blob 65,697 B) through
GpuCodec.pack_batch / unpack_batch (host buffers, Python binding included) and, for reference, hashlib+gzip on one core
(bounded sample).  Run on a GPU box: python tools/c3_small_blobs.py [n_blobs]"""
import gzip, hashlib, os, pickle, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metaflow_b200.codec import GpuCodec

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rng = np.random.default_rng(5)
base = rng.standard_normal((512, 16384), dtype=np.float32)
blobs = []
for i in range(n):
    a = base[i % 512].copy(); a[0] = float(i)                 # distinct content -> distinct keys
    blobs.append(pickle.dumps(a, protocol=4))
nb = sum(len(b) for b in blobs)
codec = GpuCodec(devices=[0], wave_bytes=1 << 30)
for attempt in range(2):
    t0 = time.perf_counter(); keys, packed = codec.pack_batch(blobs); t1 = time.perf_counter()
    pk = [bytes(p) for p in packed]
    t2 = time.perf_counter(); back = codec.unpack_batch(pk, as_bytes=False); t3 = time.perf_counter()
    print("pass %d: %d blobs, %.2f GB: pack %.3f s (%.2f GB/s, %.0f blobs/s)  unpack %.3f s (%.2f GB/s)  ratio %.4f" % (
        attempt, n, nb / 1e9, t1 - t0, nb / (t1 - t0) / 1e9, n / (t1 - t0), t3 - t2, nb / (t3 - t2) / 1e9, sum(map(len, pk)) / nb))
for i in (0, 1, n // 2, n - 1):
    assert keys[i] == hashlib.sha1(blobs[i]).hexdigest() and gzip.decompress(pk[i]) == blobs[i] and bytes(back[i]) == blobs[i]
assert len(set(keys)) == n
m = 512
t0 = time.perf_counter()
for b in blobs[:m]:
    hashlib.sha1(b).hexdigest(); z = gzip.compress(b, 3)
t1 = time.perf_counter()
for b in blobs[:m]:
    gzip.decompress(gzip.compress(b, 3)) if False else None
print("CPU one core (hashlib.sha1 + gzip level 3), %d blobs: %.2f ms per blob -> %.3f GB/s" % (m, (t1 - t0) / m * 1e3, m * len(blobs[0]) / (t1 - t0) / 1e9))
codec.close()
