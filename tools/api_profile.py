"""Where the time of the object-level path goes (TaskDataStore.save_artifacts / load_artifacts on tmpfs).
Run on a GPU box: python tools/api_profile.py [n_blobs]"""
import os, sys, time, shutil, tempfile, functools, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metaflow_b200 import FlowDataStore, LocalStorage
from metaflow_b200 import codec as codec_mod, serializer
from metaflow_b200.datastore import task_datastore as tds_mod

acc = collections.defaultdict(float)
def timed(name, fn):
    @functools.wraps(fn)
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] += time.perf_counter() - t
    return w

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
codec = codec_mod.GpuCodec(devices=[0], wave_bytes=1 << 30)
codec._stage_inputs = timed("stage_inputs", codec._stage_inputs)
lib = codec._lib
class LibProxy(object):
    def __init__(self, lib): self._l = lib
    def __getattr__(self, k):
        v = getattr(self._l, k)
        if k in ("mfcas_pack_batch_extents", "mfcas_unpack_batch", "mfcas_hash_batch"):
            return timed(k, v)
        return v
codec._lib = LibProxy(lib)
LocalStorage.save_views = timed("save_views", LocalStorage.save_views)
tds_mod.dumps_segments = timed("pickle(dumps_segments)", serializer.dumps_segments)
import pickle
class P(object):
    loads = staticmethod(timed("pickle.loads", pickle.loads)); dumps = staticmethod(pickle.dumps)
    def __getattr__(self, k): return getattr(pickle, k)
tds_mod.pickle = P()

arrays = [np.random.default_rng(i).standard_normal(4194304, dtype=np.float32) for i in range(n)]
nb = n * (arrays[0].nbytes + 163)
root = tempfile.mkdtemp(prefix="mfcas_prof_", dir="/dev/shm")
try:
    fds = FlowDataStore("ProfFlow", storage_impl=LocalStorage, ds_root=root, codec=codec)
    for attempt in range(3):
        acc.clear()
        shutil.rmtree(os.path.join(root, "ProfFlow", "data"), ignore_errors=True)
        w = fds.get_task_datastore("1", "work", str(attempt), attempt=0, mode="w"); w.init_task()
        t0 = time.perf_counter()
        w.save_artifacts(((("a%d" % i), a) for i, a in enumerate(arrays)), len_hint=n); w.done()
        t1 = time.perf_counter()
        save_acc = dict(acc); acc.clear()
        r = fds.get_task_datastore("1", "work", str(attempt))
        back = dict(r.load_artifacts(["a%d" % i for i in range(n)]))
        t2 = time.perf_counter()
        load_acc = dict(acc)
        print("pass %d: save %.3f s (%.2f GB/s)  load %.3f s (%.2f GB/s)" % (attempt, t1 - t0, nb / (t1 - t0) / 1e9, t2 - t1, nb / (t2 - t1) / 1e9))
        print("   save:", {k: round(v, 3) for k, v in save_acc.items()})
        print("   load:", {k: round(v, 3) for k, v in load_acc.items()})
    assert all(np.array_equal(back["a%d" % i], a) for i, a in enumerate(arrays))
finally:
    codec.close(); shutil.rmtree(root, ignore_errors=True)
