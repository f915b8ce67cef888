"""Host logic of the datastore classes, CPU only: the GPU codec is replaced by the test-only
OracleCodec so that batching, dedup, file layout, metadata files and error text can be checked
against the golden fixtures and -- when /root/reference exists (build container) -- against the
real reference reading our files and vice versa."""
import base64
import json
import os
import sys

import numpy as np
import pytest

import golden_inputs
from oracle_codec import OracleCodec
from metaflow_b200 import FlowDataStore, LocalStorage, TaskDataStoreSet
from metaflow_b200.datastore.content_addressed_store import ContentAddressedStore
from metaflow_b200.tuple_util import ForeachFrame
from metaflow_b200.exception import DataException, MetaflowInternalError, UnpicklableArtifactException

HAVE_REF = os.path.isdir("/root/reference/metaflow")


def _eq(a, b):
    if isinstance(a, np.ndarray):
        return isinstance(b, np.ndarray) and a.dtype == b.dtype and np.array_equal(a, b)
    return a == b


@pytest.fixture
def fds(tmp_path):
    return FlowDataStore("GoldenFlow", storage_impl=LocalStorage, ds_root=str(tmp_path), codec=OracleCodec())


def test_save_blobs_keys_layout_and_order(fds, tmp_path, golden):
    blobs = golden_inputs.raw_blobs()
    names = list(blobs)
    res = fds.ca_store.save_blobs(iter(blobs[n] for n in names), len_hint=len(names))
    assert [r.key for r in res] == [golden["raw_blobs"][n]["key"] for n in names]
    assert all(r.uri is None for r in res)
    for n, r in zip(names, res):
        path = tmp_path / "GoldenFlow" / "data" / r.key[:2] / r.key
        assert path.is_file()
        assert json.loads((tmp_path / "GoldenFlow" / "data" / r.key[:2] / (r.key + "_meta")).read_text()) == {"cas_raw": False, "cas_version": 1}
    got = dict(fds.ca_store.load_blobs([r.key for r in res]))
    for n, r in zip(names, res):
        assert got[r.key] == blobs[n]
    assert not [p for p in (tmp_path / "GoldenFlow" / "data").rglob("tmp*")]   # temp files were renamed away


def test_one_gpu_batch_per_window_and_dedup(fds, tmp_path):
    codec = fds.ca_store._codec
    blobs = [b"same"] * 5 + [b"other"]
    res = fds.ca_store.save_blobs(iter(blobs))
    assert codec.calls["pack"] == 1                      # whole iterator -> one batch
    assert len({r.key for r in res}) == 2 and [r.key for r in res][:5] == [res[0].key] * 5
    before = {p: p.stat().st_mtime_ns for p in (tmp_path / "GoldenFlow" / "data").rglob("*") if p.is_file()}
    res_again = fds.ca_store.save_blobs(iter(blobs))     # existing keys are never rewritten ...
    after = {p: p.stat().st_mtime_ns for p in (tmp_path / "GoldenFlow" / "data").rglob("*") if p.is_file()}
    assert before == after and res_again == res
    assert codec.calls["pack"] == 1                      # ... and tiny blobs keyed before cost no codec call at all
    # dedup happens BEFORE pack: only the two distinct blobs ever reached the codec
    assert codec.packed_blobs == 2
    # a key remembered from earlier whose file has vanished is packed and written again
    victim = tmp_path / "GoldenFlow" / "data" / res[0].key[:2] / res[0].key
    victim.unlink()
    assert fds.ca_store.save_blobs(iter([b"same"]))[0].key == res[0].key and victim.is_file()
    assert codec.calls["pack"] == 2
    fds.ca_store._window_bytes = 8                       # small window -> several batches, same result
    res2 = fds.ca_store.save_blobs(iter([b"0123456789", b"abcdefghij", b"xyz"]))
    assert codec.calls["pack"] == 2 + 3
    assert [r.key for r in res2] == codec.hash_batch([b"0123456789", b"abcdefghij", b"xyz"])


def test_raw_blobs_and_load_data(fds):
    out = fds.save_data(iter([b"package-bytes", b""]))
    assert out[0][0].endswith(out[0][1]) and out[0][0].startswith(fds.datastore_root)
    back = dict(fds.load_data([k for _, k in out]))
    assert back[out[0][1]] == b"package-bytes" and back[out[1][1]] == b""
    with open(out[0][0], "rb") as f:
        assert f.read() == b"package-bytes"               # raw: the URI holds the bytes verbatim


def test_task_datastore_matches_reference_data_json(fds, tmp_path, golden):
    objs = golden_inputs.artifact_objects()
    w = fds.get_task_datastore("1", "start", "7", attempt=0, mode="w")
    w.init_task()
    w.save_artifacts(iter(objs.items()), len_hint=len(objs))
    w.done()
    tdir = tmp_path / "GoldenFlow" / "1" / "start" / "7"
    data_json = json.loads((tdir / "0.data.json").read_text())
    for k in ("datastore", "version", "attempt", "objects", "info"):
        assert data_json[k] == golden["data_json"][k], k
    assert data_json["python_version"] == sys.version
    assert (tdir / "0.DONE.lock").read_bytes() == base64.b64decode(golden["done_lock_b64"])
    assert "time" in json.loads((tdir / "0.attempt.json").read_text())
    r = fds.get_task_datastore("1", "start", "7")
    assert r.attempt == 0
    back = dict(r.load_artifacts(list(objs)))
    for n, o in objs.items():
        assert _eq(o, back[n]), n
    assert r["answer"] == 42 and "answer" in r and r.get("nope", 5) == 5 and r.is_none("none")
    assert dict(r.get_artifact_sizes(["answer"]))["answer"] == golden["data_json"]["info"]["answer"]["size"]
    a, b = (x[1] for x in r.load_artifacts(["true", "dup_of_true"]))
    assert a is True and b is True
    with pytest.raises(MetaflowInternalError):
        w.save_artifacts(iter([("x", 1)]))                # write after done()
    with pytest.raises(MetaflowInternalError):
        r.save_artifacts(iter([("x", 1)]))                # wrong mode


def test_unpicklable_and_missing_attempts(fds):
    w = fds.get_task_datastore("1", "s", "1", attempt=0, mode="w")
    w.init_task()
    with pytest.raises(UnpicklableArtifactException) as e:
        w.save_artifacts(iter([("gen", (i for i in range(3)))]))
    assert 'Cannot pickle dump artifact named "gen"' in str(e.value)
    with pytest.raises(DataException) as e:
        fds.get_task_datastore("1", "s", "1")             # started, never finished
    assert "No completed attempts of the task was found" in str(e.value)
    assert fds.get_task_datastore("1", "s", "1", allow_not_done=True).attempt == 0


def test_attempt_resolution_and_datastore_set(fds):
    for task, val in (("10", "a"), ("11", "b"), ("12", "c")):
        for attempt in (0, 1):
            w = fds.get_task_datastore("5", "inner", task, attempt=attempt, mode="w")
            w.init_task()
            w.save_artifacts(iter([("x", val + str(attempt)), ("shared", "same-for-all"),
                                   ("_foreach_stack", [ForeachFrame("outer", "v", 3, int(task) - 10, val)])]))
            if not (task == "12" and attempt == 1):
                w.done()
    dss = fds.get_task_datastores(pathspecs=["5/inner/11", "5/inner/10"])
    assert [d.task_id for d in dss] == ["11", "10"]       # caller order preserved (foreach join inputs)
    assert [d["x"] for d in dss] == ["b1", "a1"] and all(d.attempt == 1 for d in dss)
    all_done = fds.get_task_datastores(run_id="5", steps=["inner"])
    assert sorted(d.task_id for d in all_done) == ["10", "11"]   # task 12's latest attempt never finished
    codec = fds.ca_store._codec
    n0 = codec.calls["unpack"]
    s = TaskDataStoreSet(fds, "5", steps=["inner"], prefetch_data_artifacts=["shared", "x", "_foreach_stack"])
    assert codec.calls["unpack"] == n0 + 1                # ONE batched unpack for the whole fan-in
    n1 = codec.calls["unpack"]
    assert sorted(d["shared"] for d in s) == ["same-for-all"] * 2
    assert s.get_with_pathspec_index("5/inner[1]").task_id == "11" and s.get_with_pathspec("5/inner/10")["x"] == "a1"
    assert codec.calls["unpack"] == n1                    # served from the ImmutableBlobCache
    fds.ca_store.set_blob_cache(None)


def test_passdown_clone_and_persist(fds):
    class FakeFlow(object):
        _EPHEMERAL = {"_EPHEMERAL", "_datastore"}
        _datastore = None

        def method(self):
            return 1

    o = fds.get_task_datastore("2", "a", "1", attempt=0, mode="w")
    o.init_task()
    flow = FakeFlow()
    flow.x = np.arange(10)
    flow.name = "keepme"
    flow._transition = (["b"], None)
    o.persist(flow)
    o.done()
    assert not hasattr(flow, "x") and flow.name == "keepme"      # destructive for user artifacts only
    origin = fds.get_task_datastore("2", "a", "1")
    assert set(origin) == {"x", "name", "_transition"}
    n = fds.get_task_datastore("2", "b", "2", attempt=0, mode="w")
    n.init_task()
    n.passdown_partial(origin, ["x", "missing"])
    n.done()
    assert fds.get_task_datastore("2", "b", "2").keys_for_artifacts(["x"]) == origin.keys_for_artifacts(["x"])
    c = fds.get_task_datastore("2", "c", "3", attempt=0, mode="w")
    c.init_task()
    c.clone(origin)
    c.done()
    assert np.array_equal(fds.get_task_datastore("2", "c", "3")["x"], np.arange(10))


def test_logs_and_metadata_roundtrip(fds):
    w = fds.get_task_datastore("3", "s", "1", attempt=0, mode="w")
    w.init_task()
    w.save_metadata({"task_begin": {"a": 1}})
    w.save_logs("runtime", {"stdout": b"hello\n", "stderr": b""})
    w.done()
    r = fds.get_task_datastore("3", "s", "1")
    assert r.load_metadata(["task_begin", "nope"]) == {"task_begin": {"a": 1}, "nope": None}
    assert r.load_logs(["runtime"], "stdout") == [("runtime", b"hello\n")]
    assert r.get_log_size(["runtime"], "stdout") == 6
    assert r.get_log_location("runtime", "stdout").endswith("3/s/1/0.runtime_stdout.log")


# --- the reference's own unit test for load_blobs error messages, re-expressed against our class
# (test/unit/test_content_addressed_store.py:49-95)
class _FakeStorage(object):
    TYPE = "fake"

    def __init__(self, entries):
        self._entries = entries

    @staticmethod
    def path_join(*parts):
        return "/".join(parts)

    @staticmethod
    def path_split(path):
        return path.split("/")

    @staticmethod
    def full_uri(path):
        return "fake://" + path

    def load_bytes(self, paths):
        from contextlib import contextmanager

        @contextmanager
        def cm():
            yield iter(self._entries)

        assert {e[0] for e in self._entries}.issubset(set(paths))
        return cm()


@pytest.mark.parametrize("meta, unpack_error, expected", [
    ({}, None, ["Could not extract encoding version"]),
    ({"cas_version": 999, "cas_raw": False}, None, ["Unknown encoding version 999"]),
    ({"cas_version": 1, "cas_raw": False}, "boom", ["Could not unpack artifact", "boom"]),
    ({"cas_version": 1, "cas_raw": False}, "codec", ["Could not unpack artifact"]),
])
def test_load_blobs_error_messages(tmp_path, monkeypatch, meta, unpack_error, expected):
    current_path, stale_path = "prefix/bb/bbbbbbbbbb", "prefix/aa/aaaaaaaaaa"
    f = tmp_path / "blob.bin"
    f.write_bytes(b"not-a-valid-gzip-stream")
    store = ContentAddressedStore("prefix", _FakeStorage([(current_path, str(f), meta)]), codec=OracleCodec())
    if unpack_error == "boom":
        def _raise(_fileobj):
            raise ValueError("boom")
        monkeypatch.setattr(store, "_unpack_v1", _raise)
    with pytest.raises(DataException) as exc:
        list(store.load_blobs(["bbbbbbbbbb", "aaaaaaaaaa"]))
    msg = str(exc.value)
    assert current_path in msg and stale_path not in msg
    for e in expected:
        assert e in msg


# --- cross-implementation checks with the REAL reference (build container only)
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")
def test_reference_reads_ours_and_we_read_reference(tmp_path):
    import subprocess
    import textwrap

    objs = golden_inputs.artifact_objects()
    ours = FlowDataStore("XFlow", storage_impl=LocalStorage, ds_root=str(tmp_path / "ours"), codec=OracleCodec())
    w = ours.get_task_datastore("1", "start", "1", attempt=0, mode="w")
    w.init_task()
    w.save_artifacts(iter(objs.items()))
    w.done()
    script = textwrap.dedent("""
        import sys, os, json
        sys.path.insert(0, "/root/reference"); sys.path.insert(0, %r)
        os.environ.setdefault("USER", "x")
        import numpy as np, golden_inputs
        from metaflow.datastore import FlowDataStore
        from metaflow.plugins.datastores.local_storage import LocalStorage
        objs = golden_inputs.artifact_objects()
        r = FlowDataStore("XFlow", storage_impl=LocalStorage, ds_root=%r).get_task_datastore("1", "start", "1")
        back = dict(r.load_artifacts(list(objs)))
        for n, o in objs.items():
            assert (np.array_equal(o, back[n]) if isinstance(o, np.ndarray) else o == back[n]), n
        f = FlowDataStore("XFlow", storage_impl=LocalStorage, ds_root=%r)
        w = f.get_task_datastore("1", "start", "1", attempt=0, mode="w"); w.init_task()
        w.save_artifacts(iter(objs.items())); w.done()
        print("REF-OK")
    """) % (os.path.dirname(os.path.abspath(__file__)), str(tmp_path / "ours"), str(tmp_path / "theirs"))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONPATH", None)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert "REF-OK" in out.stdout, out.stderr[-2000:]
    theirs = FlowDataStore("XFlow", storage_impl=LocalStorage, ds_root=str(tmp_path / "theirs"), codec=OracleCodec())
    r = theirs.get_task_datastore("1", "start", "1")
    back = dict(r.load_artifacts(list(objs)))
    for n, o in objs.items():
        assert _eq(o, back[n]), n
    a = json.loads((tmp_path / "ours" / "XFlow" / "1" / "start" / "1" / "0.data.json").read_text())
    b = json.loads((tmp_path / "theirs" / "XFlow" / "1" / "start" / "1" / "0.data.json").read_text())
    assert a == b
