"""GPU parity tests at the BASELINE.json SHAPES (run with -m gpu on a B200), through the product's public layers
(FlowRunner / TaskDataStore / TaskDataStoreSet / ContentAddressedStore -> GpuCodec -> C ABI), against the CPU oracle:

  C1  FlowSpec foreach x8, 1 MiB artifacts, join            -> FlowRunner with GpuCodec, keys == the reference's
  --  nested foreach + joins + merge_artifacts + persist()   -> same, through the runner (rows a2, a13, a14, n3)
  C3  foreach x65536, 64 KiB artifacts (blob 65,697 B)       -> all 65,536 keys == hashlib, all blobs round trip
  C4  nested foreach 256 x 256, 1 MiB artifacts, join loads  -> every inner join = ONE batched unpack of 256 inputs,
                                                                on blobs we wrote and on reference-written blobs
  --  16 MiB artifacts written by the REAL reference         -> tests/golden/ref_written_16mib.json decodes on the GPU
  --  one artifact > 4 GiB (test/core/tests/large_artifact.py:29)
  --  forked task processes share the parent's GPU context   -> multicore_utils.parallel_map + the batcher

Everything big lives under /dev/shm (tmpfs) and is removed afterwards.
"""
import base64
import gzip
import hashlib
import json
import os
import pickle
import shutil
import tempfile
import textwrap
import time

import numpy as np
import pytest

import golden_inputs
from oracle import c_oracle, ref_path

pytestmark = pytest.mark.gpu


@pytest.fixture()
def shm_dir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="mfcas_shape_", dir=base)
    yield d
    shutil.rmtree(d, ignore_errors=True)


def _free_bytes(path):
    st = os.statvfs(path)
    return st.f_bavail * st.f_frsize


C1_FLOW = textwrap.dedent('''
    import numpy as np
    from metaflow_b200.flowspec import FlowSpec, step

    class ForeachFlow(FlowSpec):
        @step
        def start(self):
            self.seeds = list(range(8))
            self.common = "shared-by-all"
            self.next(self.work, foreach="seeds")

        @step
        def work(self):
            self.arr = np.random.default_rng(self.input).standard_normal(262144, dtype=np.float32)   # 1 MiB
            self.seed = self.input
            self.next(self.join)

        @step
        def join(self, inputs):
            self.total = float(sum(float(inp.arr.sum()) for inp in inputs))
            self.order = [inp.seed for inp in inputs]
            self.merge_artifacts(inputs, exclude=["arr", "seed"])
            self.next(self.end)

        @step
        def end(self):
            self.done_flag = True
''')

NESTED_FLOW = textwrap.dedent('''
    import numpy as np
    from metaflow_b200.flowspec import FlowSpec, step

    class NestedFlow(FlowSpec):
        @step
        def start(self):
            self.outer = list(range(OUTER))
            self.tag = "nested"
            self.next(self.fan, foreach="outer")

        @step
        def fan(self):
            self.o = self.input
            self.inner = [self.input * 1000 + j for j in range(INNER)]
            self.next(self.leaf, foreach="inner")

        @step
        def leaf(self):
            self.x = np.random.default_rng(self.input).standard_normal(ELEMS, dtype=np.float32)
            self.ident = self.input
            self.next(self.inner_join)

        @step
        def inner_join(self, inputs):
            self.sums = [float(i.x.sum(dtype=np.float64)) for i in inputs]
            self.idents = [i.ident for i in inputs]
            self.merge_artifacts(inputs, include=["o", "tag"])
            self.next(self.outer_join)

        @step
        def outer_join(self, inputs):
            self.all_sums = [s for i in inputs for s in i.sums]
            self.all_idents = [k for i in inputs for k in i.idents]
            self.merge_artifacts(inputs, include=["tag"])
            self.next(self.end)

        @step
        def end(self):
            self.n = len(self.all_idents)
''')


def _flow(src, name, **consts):
    ns = dict(consts)
    exec(compile(src, name + ".py", "exec"), ns)
    return ns[name]


def _assert_reference_can_read(root, flow_name, ds, names):
    """Every stored file of `names` is plain gzip to the reference's reader, and the keys are the reference's keys."""
    for n in names:
        key = ds._objects[n]
        z = open(os.path.join(root, flow_name, "data", key[:2], key), "rb").read()
        blob = ref_path.unpack_v1(z)
        assert ref_path.blob_key(blob) == key, n
        e, out = c_oracle.gunzip(z, len(blob))
        assert e == 0 and out == blob, n
        meta = json.load(open(os.path.join(root, flow_name, "data", key[:2], key + "_meta")))
        assert meta == {"cas_raw": False, "cas_version": 1}


def test_c1_flow_through_the_runner_on_gpu(gpu_codec, tmp_path):
    """BASELINE config C1 on the GPU path: FlowSpec foreach x8 of 1 MiB artifacts + join, in-process runner, GpuCodec."""
    from metaflow_b200.flowspec import FlowRunner
    root = str(tmp_path / "ds")
    l0 = gpu_codec.counters()["kernel_launches"]
    run = FlowRunner(_flow(C1_FLOW, "ForeachFlow"), ds_root=root, codec=gpu_codec).run()
    assert gpu_codec.counters()["kernel_launches"] > l0
    work = run.task_datastores("work")
    want = [np.random.default_rng(i).standard_normal(262144, dtype=np.float32) for i in range(8)]
    # keys are what the reference computes for the same objects (oracle: pickle.dumps + hashlib)
    ref_objects, ref_info = ref_path.save_artifacts(str(tmp_path / "ref"), "ForeachFlow", [("arr", a) for a in want][:1])
    for i, ds in enumerate(work):
        assert ds._objects["arr"] == hashlib.sha1(pickle.dumps(want[i], protocol=4)).hexdigest()
        assert ds._info["arr"] == {"size": 1048739, "type": "<class 'numpy.ndarray'>", "encoding": "gzip+pickle-v4"}
        assert ds["seed"] == i and np.array_equal(ds["arr"], want[i]) and ds["common"] == "shared-by-all"
        assert ds.pathspec_index == "1/work[%d]" % i
        _assert_reference_can_read(root, "ForeachFlow", ds, ["arr", "seed", "_foreach_stack", "name"])
    assert work[0]._objects["arr"] == ref_objects["arr"] and work[0]._info["arr"] == ref_info["arr"]
    j = run.task_datastores("join")[0]
    assert j["order"] == list(range(8)) and abs(j["total"] - float(sum(float(a.sum()) for a in want))) < 1e-3
    assert j["common"] == "shared-by-all" and "arr" not in j
    assert run.task_datastores("end")[0]["done_flag"] is True
    dj = json.load(open(os.path.join(root, "ForeachFlow", "1", "work", work[3].task_id, "0.data.json")))
    assert dj["version"] == "1.0" and dj["datastore"] == "local"


def test_nested_foreach_joins_merge_and_persist_on_gpu(gpu_codec, tmp_path):
    """Rows a2 / a13 / a14 / n3 on hardware: nested foreach 4 x 6 with two join levels and merge_artifacts through the
    runner with GpuCodec; every inner join loads its inputs' arrays with ONE batched unpack; TaskDataStore.persist()
    (the reference's entry, task_datastore.py:744-809) saves a flow object's attributes on the GPU."""
    from metaflow_b200 import FlowDataStore, LocalStorage
    from metaflow_b200.flowspec import FlowRunner, FlowSpec
    root = str(tmp_path / "ds")
    cls = _flow(NESTED_FLOW, "NestedFlow", OUTER=4, INNER=6, ELEMS=40000)
    run = FlowRunner(cls, ds_root=root, codec=gpu_codec).run()
    end = run.task_datastores("end")[0]
    idents = [o * 1000 + j for o in range(4) for j in range(6)]
    assert end["n"] == 24 and end["all_idents"] == idents and end["tag"] == "nested"
    want = [float(np.random.default_rng(k).standard_normal(40000, dtype=np.float32).sum(dtype=np.float64)) for k in idents]
    assert np.allclose(end["all_sums"], want, rtol=0, atol=1e-6)
    leaves = run.task_datastores("leaf")
    assert len(leaves) == 24
    for ds in leaves:
        k = ds["ident"]
        stack = ds["_foreach_stack"]
        assert [f.index for f in stack] == [k // 1000, k % 1000] and [f.num_splits for f in stack] == [4, 6]
        assert ds._objects["x"] == hashlib.sha1(pickle.dumps(np.random.default_rng(k).standard_normal(40000, dtype=np.float32), protocol=4)).hexdigest()
    _assert_reference_can_read(root, "NestedFlow", leaves[7], ["x", "ident", "_foreach_stack"])
    ij = run.task_datastores("inner_join")
    assert len(ij) == 4 and sorted(d["o"] for d in ij) == [0, 1, 2, 3]
    # persist(): attributes of a flow object -> artifacts, destructive hand-over, same keys as the oracle
    fds = FlowDataStore("PersistFlow", storage_impl=LocalStorage, ds_root=root, codec=gpu_codec)

    class PersistFlow(FlowSpec):
        pass

    f = PersistFlow()
    f.weights = np.arange(300000, dtype=np.float32)
    f.note = "persisted"
    w = fds.get_task_datastore("9", "start", "1", attempt=0, mode="w")
    w.init_task()
    w.persist(f)
    w.done()
    assert "weights" not in f.__dict__
    r = fds.get_task_datastore("9", "start", "1")
    assert r._objects["weights"] == hashlib.sha1(pickle.dumps(np.arange(300000, dtype=np.float32), protocol=4)).hexdigest()
    assert r["note"] == "persisted" and np.array_equal(r["weights"], np.arange(300000, dtype=np.float32))


def test_c3_full_shape_65536_small_blobs(gpu_codec):
    """BASELINE config C3 at shape: foreach x65536 of 64 KiB fp32 artifacts (pickle blob 65,697 B, 4.31 GB in total).
    Every key is checked against hashlib, every blob round-trips through the GPU decoder, and a 1-in-257 sample of the
    emitted files is decoded by the reference's reader (CPython gzip) and by the independent C inflater."""
    n = 65536
    rng = np.random.default_rng(5)
    base = rng.standard_normal((512, 16384), dtype=np.float32)
    blobs = []
    for i in range(n):
        a = base[i % 512].copy()
        a[0] = float(i)                               # distinct content -> distinct keys (dedup must not short-circuit)
        blobs.append(pickle.dumps(a, protocol=4))
    assert len(blobs[0]) == 65697 and sum(map(len, blobs)) == n * 65697
    t0 = time.perf_counter()
    keys, packed = gpu_codec.pack_batch(blobs)
    t1 = time.perf_counter()
    assert len(set(keys)) == n
    assert keys == [hashlib.sha1(b).hexdigest() for b in blobs]
    files = [bytes(p) for p in packed]
    for i in range(0, n, 257):
        assert gzip.decompress(files[i]) == blobs[i], i
        e, out = c_oracle.gunzip(files[i], len(blobs[i]))
        assert e == 0 and out == blobs[i], i
    t2 = time.perf_counter()
    back = gpu_codec.unpack_batch(files, as_bytes=False)
    t3 = time.perf_counter()
    agg, want = hashlib.sha1(), hashlib.sha1()
    for b, o in zip(blobs, back):
        assert not isinstance(o, Exception)
        agg.update(o)
        want.update(b)
    assert agg.digest() == want.digest()                # checksum of everything decoded == checksum of the inputs
    nb = n * 65697
    print("C3 at shape: pack %.2f s (%.2f GB/s, %.0f blobs/s), unpack %.2f s (%.2f GB/s), ratio %.4f"
          % (t1 - t0, nb / (t1 - t0) / 1e9, n / (t1 - t0), t3 - t2, nb / (t3 - t2) / 1e9, sum(map(len, files)) / nb))


def _c4_geometry(shm_dir):
    """(outer, inner, elems): the BASELINE shape 256 x 256 x 1 MiB when tmpfs can hold it (64 GiB packed + slack)."""
    if _free_bytes(shm_dir) > (110 << 30) and not os.environ.get("MFCAS_TEST_SMALL"):
        return 256, 256, 262144
    return 256, 256, 16384


def test_c4_nested_256x256_join_loads(gpu_codec, shm_dir):
    """BASELINE config C4 at shape: 65,536 leaf artifacts of a nested 256 x 256 foreach (1 MiB fp32 each, blob
    1,048,739 B, 68.7 GB in total) are saved the way the runner saves a foreach step -- the 256 sibling tasks of an
    inner foreach as ONE ContentAddressedStore.save_blobs batch, each task its own TaskDataStore -- then each of the 256 inner joins
    opens its 256 inputs with TaskDataStoreSet and prefetches their arrays -- ONE batched GPU unpack per join -- and
    the outer join loads the 256 inner results.  Checked: every key == hashlib of the reference's pickle, a sample of
    files decodes with the reference's reader, every loaded array equals its input (sum of all elements per join +
    full comparison of sampled arrays)."""
    from metaflow_b200 import FlowDataStore, LocalStorage, TaskDataStoreSet
    from metaflow_b200.tuple_util import ForeachFrame
    outer, inner, elems = _c4_geometry(shm_dir)
    fds = FlowDataStore("C4Flow", storage_impl=LocalStorage, ds_root=shm_dir, codec=gpu_codec)
    rng = np.random.default_rng(44)
    base = rng.standard_normal((64, elems), dtype=np.float32)

    def leaf_array(o, j):
        a = base[(o * 7 + j) % 64].copy()
        a[0], a[1] = float(o), float(j)
        return a

    t_save = t_load = 0.0
    sums = np.zeros((outer, inner))
    for o in range(outer):
        arrays = [leaf_array(o, j) for j in range(inner)]
        t0 = time.perf_counter()
        # one save per inner foreach step: 256 sibling tasks, each its own task datastore, ONE batch of blobs
        stores = []
        for j in range(inner):
            w = fds.get_task_datastore("1", "leaf", "%d_%d" % (o, j), attempt=0, mode="w")
            w.init_task()
            stores.append(w)
        stacks = [[ForeachFrame("start", "outer", outer, o, o), ForeachFrame("fan", "inner", inner, j, j)] for j in range(inner)]
        blobs = [pickle.dumps(a, protocol=4) for a in arrays] + [pickle.dumps(st, protocol=4) for st in stacks]
        keys = fds.ca_store.save_blobs(iter(blobs), len_hint=len(blobs))
        for j, w in enumerate(stores):
            w._objects.update({"x": keys[j].key, "_foreach_stack": keys[inner + j].key})
            w._info["x"] = {"size": len(blobs[j]), "type": "<class 'numpy.ndarray'>", "encoding": "gzip+pickle-v4"}
            w._info["_foreach_stack"] = {"size": len(blobs[inner + j]), "type": "<class 'list'>", "encoding": "gzip+pickle-v4"}
            w.done()
        t_save += time.perf_counter() - t0
        if o % 32 == 0:
            for j in (0, inner - 1):
                assert keys[j].key == hashlib.sha1(pickle.dumps(arrays[j], protocol=4)).hexdigest()
                z = open(os.path.join(shm_dir, "C4Flow", "data", keys[j].key[:2], keys[j].key), "rb").read()
                assert ref_path.unpack_v1(z) == pickle.dumps(arrays[j], protocol=4)
        sums[o] = [float(a.sum(dtype=np.float64)) for a in arrays]
    # ---- the load side: every inner join = TaskDataStoreSet over its 256 inputs, prefetch = ONE batched unpack
    inner_results = []
    for o in range(outer):
        t0 = time.perf_counter()
        l0 = gpu_codec.counters()["kernel_launches"]
        # (the control artifact goes into the prefetch as in the reference's join, task.py:252-279: the set indexes its
        # datastores by pathspec_index, which reads every input's _foreach_stack)
        dss = TaskDataStoreSet(fds, "1", pathspecs=["1/leaf/%d_%d" % (o, j) for j in range(inner)], prefetch_data_artifacts=["x", "_foreach_stack"])
        l1 = gpu_codec.counters()["kernel_launches"]
        got = [ds["x"] for ds in dss]                             # served from the prefetch cache, pathspec order
        assert gpu_codec.counters()["kernel_launches"] == l1
        t_load += time.perf_counter() - t0
        assert len(got) == inner and 0 < l1 - l0 <= 8             # one batch (parse / inflate / verify / serial-foreign)
        assert [float(a.sum(dtype=np.float64)) for a in got] == list(sums[o])
        assert all(a[0] == float(o) and a[1] == float(j) for j, a in enumerate(got))
        if o % 64 == 0:
            assert np.array_equal(got[5], leaf_array(o, 5)) and np.array_equal(got[inner - 1], leaf_array(o, inner - 1))
        fds.ca_store.set_blob_cache(None)
        w = fds.get_task_datastore("1", "inner_join", str(o), attempt=0, mode="w")
        w.init_task()
        w.save_artifacts(iter([("sums", np.array([float(a.sum(dtype=np.float64)) for a in got])),
                               ("_foreach_stack", [ForeachFrame("start", "outer", outer, o, o)])]))
        w.done()
        inner_results.append(w)
    dss = TaskDataStoreSet(fds, "1", steps=["inner_join"], prefetch_data_artifacts=["sums", "_foreach_stack"])
    total = {ds.task_id: ds["sums"] for ds in dss}
    fds.ca_store.set_blob_cache(None)
    assert len(total) == outer and all(np.array_equal(total[str(o)], sums[o]) for o in range(outer))
    nb = outer * inner * (elems * 4 + 163)
    print("C4 at shape %dx%dx%d B: save %.1f s (%.2f GB/s), join loads %.1f s (%.2f GB/s)"
          % (outer, inner, elems * 4 + 163, t_save, nb / t_save / 1e9, t_load, nb / t_load / 1e9))


def test_c4_join_loads_reference_written_inputs(gpu_codec, shm_dir):
    """The C4 load side on blobs written by the reference's code path (zlib level 3, one stream, no index): one inner
    join's 256 inputs of 1 MiB through TaskDataStoreSet, one batched unpack."""
    from metaflow_b200 import FlowDataStore, LocalStorage, TaskDataStoreSet
    from metaflow_b200.tuple_util import ForeachFrame
    inner, elems = 256, 262144
    rng = np.random.default_rng(45)
    arrays = [rng.standard_normal(elems, dtype=np.float32) for _ in range(inner)]
    fds = FlowDataStore("C4Ref", storage_impl=LocalStorage, ds_root=shm_dir, codec=gpu_codec)
    for j, a in enumerate(arrays):
        stack = [ForeachFrame("fan", "inner", inner, j, j)]
        objects, info = ref_path.save_artifacts(shm_dir, "C4Ref", [("x", a), ("_foreach_stack", stack)])      # the reference's writer, restated
        w = fds.get_task_datastore("1", "leaf", str(j), attempt=0, mode="w")
        w.init_task()
        w._objects.update(objects); w._info.update(info)
        w.done()
    t0 = time.perf_counter()
    dss = TaskDataStoreSet(fds, "1", pathspecs=["1/leaf/%d" % j for j in range(inner)], prefetch_data_artifacts=["x", "_foreach_stack"])
    got = [ds["x"] for ds in dss]
    dt = time.perf_counter() - t0
    fds.ca_store.set_blob_cache(None)
    assert all(np.array_equal(g, a) for g, a in zip(got, arrays))
    print("C4 reference-written: 256 x 1 MiB join load %.2f s (%.2f GB/s)" % (dt, inner * elems * 4 / dt / 1e9))


def test_reference_written_16mib_golden_files(gpu_codec):
    """16 MiB artifacts written by the REAL reference (tests/golden/ref_written_16mib.json, oracle/make_golden_large.py)
    and N(0,1) artifacts of the same size written by the restated path decode on the GPU, bit for bit."""
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "ref_written_16mib.json")) as f:
        golden = json.load(f)
    objs = golden_inputs.large_artifacts()
    files, want = [], []
    for name, rec in golden["artifacts"].items():
        blob = pickle.dumps(objs[name], protocol=4)
        assert len(blob) == 16777379 == rec["info"]["size"] and hashlib.sha1(blob).hexdigest() == rec["key"]
        files.append(base64.b64decode(rec["gz_b64"]))
        want.append(blob)
    for seed in (1, 2):
        blob = pickle.dumps(np.random.default_rng(seed).standard_normal(4194304, dtype=np.float32), protocol=4)
        files.append(ref_path.pack_v1(blob))
        want.append(blob)
    t0 = time.perf_counter()
    got = gpu_codec.unpack_batch(files, as_bytes=False)
    dt = time.perf_counter() - t0
    for i, (g, w) in enumerate(zip(got, want)):
        assert not isinstance(g, Exception), (i, g)
        assert hashlib.sha1(g).digest() == hashlib.sha1(w).digest() and bytes(g[:4096]) == w[:4096] and len(g) == len(w), i
    # and our own packing of the same artifacts gives the reference's keys
    keys, _ = gpu_codec.pack_batch(want[:2])
    assert keys == [golden["artifacts"][n]["key"] for n in golden["artifacts"]]
    print("reference-written 16 MiB artifacts: %d files in %.2f s (%.1f MB/s per blob)" % (len(files), dt, 16.78 / dt))


def test_artifact_larger_than_4gib(gpu_codec):
    """/root/reference/test/core/tests/large_artifact.py:29 -- one 4.1 GiB artifact (b'x' * 4.1 GiB pickles to
    BINBYTES8): 64-bit lengths all the way, ISIZE wraps mod 2^32 per member, 2100 stripes / gzip members."""
    n = int(4.1 * (1 << 30))
    blob = pickle.dumps(b"x" * n, protocol=4)
    assert len(blob) > (1 << 32)
    t0 = time.perf_counter()
    keys, packed = gpu_codec.pack_batch([blob])
    t1 = time.perf_counter()
    assert keys[0] == hashlib.sha1(blob).hexdigest()
    z = bytes(packed[0])
    assert len(z) < len(blob) // 50
    with gzip.GzipFile(fileobj=__import__("io").BytesIO(z), mode="rb") as f:        # the reference's _unpack_v1, streamed
        h, total = hashlib.sha1(), 0
        while True:
            piece = f.read(64 << 20)
            if not piece:
                break
            h.update(piece); total += len(piece)
    assert total == len(blob) and h.hexdigest() == keys[0]
    t2 = time.perf_counter()
    back = gpu_codec.unpack_batch([z], as_bytes=False)[0]
    t3 = time.perf_counter()
    assert not isinstance(back, Exception) and len(back) == len(blob)
    assert hashlib.sha1(back).hexdigest() == keys[0]
    obj = pickle.loads(back)
    assert len(obj) == n and obj[:3] == b"xxx" and obj[-1:] == b"x"
    print("4.1 GiB artifact: pack %.1f s, unpack %.1f s, packed %d B" % (t1 - t0, t3 - t2, len(z)))
    del back, obj
    gpu_codec.trim(0)


def test_forked_tasks_share_the_parents_gpu_context(gpu_codec, tmp_path, monkeypatch):
    """Rows a12 / n3: multicore_utils.parallel_map forks one process per 'task'; none of them creates a CUDA context --
    their artifacts reach the GPU through the parent's batcher (shared memory + Unix socket), merged into wide batches."""
    from metaflow_b200 import batcher, multicore_utils
    from metaflow_b200.multicore_utils import parallel_map
    server = batcher.CodecServer(gpu_codec, linger=0.02).start()
    monkeypatch.setattr(multicore_utils, "_server", server)
    root = str(tmp_path / "ds")
    try:
        def task(i):
            from metaflow_b200 import FlowDataStore, LocalStorage
            from metaflow_b200 import codec as codec_mod
            assert isinstance(codec_mod.default_codec(), batcher.RemoteCodec)
            fds = FlowDataStore("ForkFlow", storage_impl=LocalStorage, ds_root=root)
            w = fds.get_task_datastore("1", "work", str(i), attempt=0, mode="w")
            w.init_task()
            arr = np.random.default_rng(i).standard_normal(1 << 20, dtype=np.float32)      # 4 MiB
            w.save_artifacts(iter([("arr", arr), ("idx", i)]))
            w.done()
            r = fds.get_task_datastore("1", "work", str(i))
            return i, r.keys_for_artifacts(["arr"])[0], float(r["arr"].sum(dtype=np.float64))

        l0 = gpu_codec.counters()["kernel_launches"]
        out = parallel_map(task, range(8), max_parallel=8)
        assert gpu_codec.counters()["kernel_launches"] > l0                       # the PARENT's context did the work
        for i, key, total in out:
            arr = np.random.default_rng(i).standard_normal(1 << 20, dtype=np.float32)
            assert key == hashlib.sha1(pickle.dumps(arr, protocol=4)).hexdigest()
            assert total == float(arr.sum(dtype=np.float64))
            z = open(os.path.join(root, "ForkFlow", "data", key[:2], key), "rb").read()
            assert ref_path.unpack_v1(z) == pickle.dumps(arr, protocol=4)
        assert server.requests >= 16 and server.batches["pack"] < 8 + 8                # merged, not one call per child
    finally:
        server.stop()
