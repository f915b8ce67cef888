"""Client read path: the file cache of metaflow/client/filecache.py in front of the datastore
classes -- cache layout, hit/miss behaviour, eviction, and the batched multi-task fetch.  The CPU oracle stands in
for the GPU codec (host logic only)."""
import os
import time

import numpy as np
import pytest

from metaflow_b200 import FlowDataStore, LocalStorage
from metaflow_b200.client import FileCache, filecache
from oracle_codec import OracleCodec


def _write_tasks(root, codec, n_tasks=6):
    fds = FlowDataStore("CacheFlow", storage_impl=LocalStorage, ds_root=str(root), codec=codec)
    arrays = {}
    for t in range(n_tasks):
        w = fds.get_task_datastore("7", "work", str(t), attempt=0, mode="w")
        w.init_task()
        arrays[t] = np.arange(5000, dtype=np.float32) * (t + 1)
        w.save_artifacts(iter([("x", arrays[t]), ("label", "task-%d" % t), ("shared", [1, 2, 3])]))
        w.done()
    return fds, arrays


def test_get_artifact_fills_and_uses_the_cache(tmp_path):
    codec = OracleCodec()
    _, arrays = _write_tasks(tmp_path / "ds", codec)
    fc = FileCache(cache_dir=str(tmp_path / "cache"), max_size=100, codec=codec)
    got = fc.get_artifact("local", str(tmp_path / "ds"), None, "CacheFlow", "7", "work", "2", "x")
    assert np.array_equal(got, arrays[2])
    unpacks = codec.calls["unpack"]
    # layout: <cache>/<type.root.flow>/<key[:2]>/<key>.blob
    ds_id = FileCache.flow_ds_id("local", str(tmp_path / "ds"), "CacheFlow")
    blobs = [f for _, _, fs in os.walk(tmp_path / "cache" / ds_id) for f in fs if f.endswith(".blob")]
    assert len(blobs) == 1 and len(blobs[0]) == 45
    # second read: served from the cache file, the codec is not called again -- also from a fresh FileCache (new process)
    again = FileCache(cache_dir=str(tmp_path / "cache"), max_size=100, codec=codec).get_artifact(
        "local", str(tmp_path / "ds"), None, "CacheFlow", "7", "work", "2", "x")
    assert np.array_equal(again, arrays[2]) and codec.calls["unpack"] == unpacks
    assert dict(fc.get_all_artifacts("local", str(tmp_path / "ds"), None, "CacheFlow", "7", "work", "2"))["label"] == "task-2"
    assert fc.get_artifact_size("local", str(tmp_path / "ds"), 0, "CacheFlow", "7", "work", "2", "x") > 20000


def test_task_metadata_is_cached_per_attempt(tmp_path):
    codec = OracleCodec()
    _write_tasks(tmp_path / "ds", codec, n_tasks=1)
    fc = FileCache(cache_dir=str(tmp_path / "cache"), max_size=100, codec=codec)
    fc.get_artifact_size("local", str(tmp_path / "ds"), 0, "CacheFlow", "7", "work", "0", "x")
    cached = [f for _, _, fs in os.walk(tmp_path / "cache") for f in fs if f.endswith(".cached")]
    assert len(cached) == 1
    # with the data.json gone the cached metadata still opens the task datastore
    os.remove(tmp_path / "ds" / "CacheFlow" / "7" / "work" / "0" / "0.data.json")
    assert fc.get_artifact_size("local", str(tmp_path / "ds"), 0, "CacheFlow", "7", "work", "0", "x") > 20000


def test_batched_fetch_is_one_load(tmp_path):
    codec = OracleCodec()
    _, arrays = _write_tasks(tmp_path / "ds", codec)
    fc = FileCache(cache_dir=str(tmp_path / "cache"), max_size=100, codec=codec)
    tasks = [("7", "work", str(t), None) for t in range(6)]
    before = codec.calls["unpack"]
    out = fc.get_artifacts_for_tasks("local", str(tmp_path / "ds"), "CacheFlow", tasks, ["x", "shared", "missing"])
    assert codec.calls["unpack"] == before + 1                       # all misses of all tasks in ONE batch
    for t in range(6):
        got = out[("7", "work", str(t))]
        assert np.array_equal(got["x"], arrays[t]) and got["shared"] == [1, 2, 3] and "missing" not in got
    assert out[("7", "work", "0")]["shared"] is not out[("7", "work", "1")]["shared"]     # same key, distinct objects
    out2 = fc.get_artifacts_for_tasks("local", str(tmp_path / "ds"), "CacheFlow", tasks, ["x"])
    assert codec.calls["unpack"] == before + 1 and np.array_equal(out2[("7", "work", "5")]["x"], arrays[5])


def test_eviction_oldest_first_with_quarantine(tmp_path, monkeypatch):
    fc = FileCache(cache_dir=str(tmp_path / "cache"), max_size=1)         # 1 MiB
    blob = b"x" * (400 << 10)
    paths = [str(tmp_path / "cache" / "local.root.F" / "ab" / ("%040d.blob" % i)) for i in range(4)]
    for p in paths[:3]:
        fc.create_file(p, blob)
    assert all(os.path.exists(p) for p in paths[:3])                 # over budget, but all younger than the quarantine
    monkeypatch.setattr(filecache, "NEW_FILE_QUARANTINE", 0)
    fc.create_file(paths[3], blob)
    left = [os.path.exists(p) for p in paths]
    assert left == [False, False, True, True]
    assert fc.read_file(paths[0]) is None and fc.read_file(paths[3]) == blob
    # a fresh FileCache indexes what is on disk
    fc2 = FileCache(cache_dir=str(tmp_path / "cache"), max_size=1)
    fc2.create_file(str(tmp_path / "cache" / "local.root.F" / "cd" / ("%040d.blob" % 9)), blob)
    assert sum(os.path.exists(p) for p in paths) == 1


def test_unknown_datastore_type(tmp_path):
    with pytest.raises(filecache.FileCacheException):
        FileCache(cache_dir=str(tmp_path))._get_flow_datastore("s3", "s3://bucket/x", "F")


def test_client_object_model_over_a_flow_run(tmp_path):
    """Flow / Run / Step / Task / DataArtifact / MetaflowData (client/core.py:973, 1106, 1511, 2257) over a
    datastore an in-process run left behind; Step.artifacts fetches a whole foreach step with ONE batched load."""
    import textwrap
    from oracle_codec import OracleCodec
    from metaflow_b200 import client
    from metaflow_b200.client import FileCache, Flow, Run, Step, Task
    from metaflow_b200.flowspec import FlowRunner

    src = textwrap.dedent("""
        import numpy as np
        from metaflow_b200.flowspec import FlowSpec, step
        class ClientFlow(FlowSpec):
            @step
            def start(self):
                self.items = [3, 5, 7, 11]
                self.next(self.work, foreach="items")
            @step
            def work(self):
                self.vec = np.full(5000, self.input, dtype=np.float32)
                self.next(self.join)
            @step
            def join(self, inputs):
                self.total = float(sum(float(i.vec[0]) for i in inputs))
                self.next(self.end)
            @step
            def end(self):
                self.answer = self.total * 2
    """)
    ns = {}
    exec(compile(src, "clientflow.py", "exec"), ns)
    codec = OracleCodec()
    root = str(tmp_path / "ds")
    FlowRunner(ns["ClientFlow"], ds_root=root, codec=codec).run()
    client.set_datastore_root(root)
    client.set_filecache(FileCache(cache_dir=str(tmp_path / "cache"), max_size=100, codec=codec))
    try:
        flow = Flow("ClientFlow")
        run = flow.latest_run
        assert run.id == "1" and run.finished and run.successful is False      # the in-process runner writes no _task_ok
        assert {s.id for s in run} == {"start", "work", "join", "end"}
        assert run.data.answer == 52.0 and "answer" in run.data and "nope" not in run.data
        step = Step("ClientFlow/1/work")
        tasks = step.tasks()
        assert len(tasks) == 4 and all(isinstance(t, Task) for t in tasks)
        calls0 = codec.calls["unpack"]
        got = step.artifacts(["vec"])
        assert codec.calls["unpack"] == calls0 + 1                             # four tasks, one batched decode
        assert sorted(float(v["vec"][0]) for v in got.values()) == [3.0, 5.0, 7.0, 11.0]
        t = tasks[0]
        assert t.index in (0, 1, 2, 3) and t.finished and t.current_attempt == 0
        calls1 = codec.calls["unpack"]                                         # (index read `_foreach_stack`)
        art = t["vec"]
        assert art.sha == t._meta()["objects"]["vec"] and art.size > 20000 and art.data.shape == (5000,)
        assert codec.calls["unpack"] == calls1                                 # ... served from the cache files now
        d = t.data
        assert float(d.vec[0]) in (3.0, 5.0, 7.0, 11.0) and [a.id for a in t] == sorted(a.id for a in t)
        assert Run("ClientFlow/1")["join"].task.data.preload().total == 26.0
        import pytest as _pytest
        with _pytest.raises(client.MetaflowNotFound):
            Task("ClientFlow/1/work/999")
    finally:
        client.set_datastore_root(None)
        client.set_filecache(None)
