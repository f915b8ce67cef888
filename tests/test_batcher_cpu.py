"""multicore_utils.parallel_map with the reference's signature, and the multi-process
batcher behind it (one process owns the codec; forked task processes reach it through shared memory).  CPU-only:
the served codec is the test-only oracle codec; the GPU variant is in test_gpu_shapes.py."""
import hashlib
import os
import threading

import numpy as np
import pytest

from oracle_codec import OracleCodec
from metaflow_b200 import batcher, multicore_utils
from metaflow_b200.multicore_utils import MulticoreException, parallel_imap_unordered, parallel_map


def test_parallel_map_reference_case():
    # /root/reference/test/unit/test_multicore_utils.py:4-12, verbatim expectation
    assert parallel_map(lambda s: s.upper(), ["a", "b", "c", "d", "e", "f"]) == ["A", "B", "C", "D", "E", "F"]


def test_parallel_map_order_closures_and_limits(tmp_path):
    offset = 1000                                       # closure state travels by fork, not by pickling
    out = parallel_map(lambda i: (i + offset, os.getpid()), range(23), max_parallel=4, dir=str(tmp_path))
    assert [v for v, _ in out] == [i + 1000 for i in range(23)]
    assert all(pid != os.getpid() for _, pid in out)
    assert not list(tmp_path.iterdir())                 # result files are removed
    big = parallel_map(lambda i: np.full(3_000_000, i, dtype=np.uint8), range(3), max_parallel=3)   # > 32 MB queue limits
    assert [int(a[0]) for a in big] == [0, 1, 2]
    assert sorted(parallel_imap_unordered(lambda x: x * x, iter(range(7)), max_parallel=2)) == [i * i for i in range(7)]
    assert parallel_map(lambda x: x, []) == []


def test_parallel_map_child_failure():
    def boom(i):
        if i == 3:
            raise ValueError("child 3 fails")
        return i

    with pytest.raises(MulticoreException):
        parallel_map(boom, range(6), max_parallel=2)


@pytest.fixture()
def server():
    srv = batcher.CodecServer(OracleCodec(), linger=0.05).start()
    yield srv
    srv.stop()


def test_remote_codec_round_trip_and_errors(server):
    c = batcher.RemoteCodec(server.socket_path)
    assert c.ping()["pid"] == os.getpid()
    rng = np.random.default_rng(4)
    blobs = [b"", b"x", rng.bytes(100_000), bytes(70_000), np.arange(5000, dtype=np.int64).tobytes()]
    keys, packed = c.pack_batch(blobs)
    assert keys == [hashlib.sha1(b).hexdigest() for b in blobs] == c.hash_batch(blobs)
    import gzip
    assert [gzip.decompress(bytes(p)) for p in packed] == blobs
    back = c.unpack_batch([bytes(p) for p in packed] + [b"not gzip"])
    assert back[:-1] == blobs and isinstance(back[-1], Exception)
    views = c.unpack_batch([bytes(packed[2])], as_bytes=False)
    assert bytes(views[0]) == blobs[2]
    assert c.pack_batch([]) == ([], []) and c.unpack_batch([]) == []
    c.close()


def test_concurrent_clients_are_merged_into_one_batch(server):
    """Requests that arrive together become ONE codec call (the batcher's whole point)."""
    n_clients, results = 6, {}
    barrier = threading.Barrier(n_clients)

    def client(k):
        c = batcher.RemoteCodec(server.socket_path)
        blobs = [b"client %d blob %d" % (k, j) * 50 for j in range(4)]
        barrier.wait()
        results[k] = (blobs, c.pack_batch(blobs)[0])
        c.close()

    threads = [threading.Thread(target=client, args=(k,)) for k in range(n_clients)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for blobs, keys in results.values():
        assert keys == [hashlib.sha1(b).hexdigest() for b in blobs]
    assert server.requests == n_clients and server.batches["pack"] < n_clients


def test_forked_tasks_save_through_the_parents_batcher(server, tmp_path, monkeypatch):
    """parallel_map children (one per 'task') persist artifacts with the ordinary datastore classes; their codec
    is the parent's, reached through the batcher -- no child creates a codec of its own."""
    monkeypatch.setattr(multicore_utils, "_server", server)
    root = str(tmp_path / "ds")

    def task(i):
        from metaflow_b200 import FlowDataStore, LocalStorage
        from metaflow_b200 import codec as codec_mod

        assert isinstance(codec_mod.default_codec(), batcher.RemoteCodec)
        fds = FlowDataStore("ForkFlow", storage_impl=LocalStorage, ds_root=root)
        w = fds.get_task_datastore("1", "work", str(i), attempt=0, mode="w")
        w.init_task()
        arr = np.random.default_rng(i).standard_normal(50_000, dtype=np.float32)
        w.save_artifacts(iter([("arr", arr), ("idx", i), ("flag", True)]))
        w.done()
        r = fds.get_task_datastore("1", "work", str(i))
        return i, r.keys_for_artifacts(["arr"])[0], float(r["arr"].sum())

    out = parallel_map(task, range(5), max_parallel=5)
    import pickle
    for i, key, total in out:
        arr = np.random.default_rng(i).standard_normal(50_000, dtype=np.float32)
        assert key == hashlib.sha1(pickle.dumps(arr, protocol=4)).hexdigest()
        assert abs(total - float(arr.sum())) < 1e-3
    assert server.requests >= 10                       # every child's save and load went through the parent
    # the parent reads what the children wrote, with its own (direct) codec
    from metaflow_b200 import FlowDataStore, LocalStorage
    fds = FlowDataStore("ForkFlow", storage_impl=LocalStorage, ds_root=root, codec=server.codec)
    assert fds.get_task_datastore("1", "work", "3")["idx"] == 3


def test_transport_leaves_nothing_behind(server):
    """Segments are anonymous memory passed by descriptor: no names in /dev/shm, no descriptor growth over many calls on either
    side (client and server share this process here), and the socket is the owner's alone."""
    import gc
    import stat

    assert stat.S_IMODE(os.stat(server.socket_path).st_mode) == 0o600
    c = batcher.RemoteCodec(server.socket_path)
    rng = np.random.default_rng(9)
    blobs = [rng.bytes(20_000), b"", b"tiny"]

    def round_trip():
        keys, packed = c.pack_batch(blobs)
        back = c.unpack_batch([bytes(p) for p in packed] + [b"garbage"])
        assert back[:3] == blobs and isinstance(back[3], Exception)
        assert c.hash_batch(blobs) == keys

    for _ in range(3):
        round_trip()
    gc.collect()
    before = len(os.listdir("/proc/self/fd"))
    for _ in range(40):
        round_trip()
    gc.collect()
    after = len(os.listdir("/proc/self/fd"))
    assert after <= before + 2, (before, after)
    if os.path.isdir("/dev/shm"):
        assert not [f for f in os.listdir("/dev/shm") if f.startswith("mfcas_seg")]
    c.close()


def test_server_survives_a_malformed_client(server):
    """Garbage on the socket (not JSON, absurd length) ends that connection, not the server."""
    import socket as _socket
    import struct as _struct

    for payload in (_struct.pack("<Q", 5) + b"\x80\x04K\x01.", _struct.pack("<Q", 1 << 60), b"\x01\x02"):
        s = _socket.socket(_socket.AF_UNIX, _socket.SOCK_STREAM)
        s.connect(server.socket_path)
        s.sendall(payload)
        s.close()
    c = batcher.RemoteCodec(server.socket_path)
    assert c.ping()["ok"] and c.hash_batch([b"abc"]) == [hashlib.sha1(b"abc").hexdigest()]
    c.close()


def test_abandoned_iteration_reaps_its_children(tmp_path):
    """A consumer that stops early leaves no zombie children and no result files (the generator's finally waits for them)."""
    it = parallel_imap_unordered(lambda i: i, range(8), max_parallel=4, dir=str(tmp_path))
    first = next(it)
    assert first in range(8)
    it.close()
    assert not list(tmp_path.iterdir())
    with pytest.raises(ChildProcessError):
        os.waitpid(-1, os.WNOHANG)                      # no child of this process is left to wait for
