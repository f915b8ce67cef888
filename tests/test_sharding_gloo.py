"""N > 1 host logic on CPU: two ranks (gloo, 127.0.0.1) shard a batch round-robin, pack their shards with the
test-only CPU codec, gather keys, and reduce a timing with MAX -- the same collectives bench.py uses under
torchrun with NCCL (which carry only these few bytes; the data path has no collective)."""
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metaflow_b200.batcher import merge_shard_results, pack_sharded, shard_indices


def _blobs(n):
    rng = np.random.default_rng(3)
    return [rng.bytes(int(rng.integers(0, 5000))) for _ in range(n)]


def _worker(rank, world, port, n, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_codec import OracleCodec
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blobs = _blobs(n)

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    keys, packed = pack_sharded(OracleCodec(), blobs, rank, world, all_gather)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)          # "my shard took 10+rank ms"
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, keys, len(packed), float(t)))
    dist.destroy_process_group()


def test_shard_indices_partition():
    for n in (0, 1, 7, 8, 33):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert seen == list(range(n))
            assert all(i % world == r for r in range(world) for i in shard_indices(n, r, world))
    assert merge_shard_results(5, [["a", "c", "e"], ["b", "d"]]) == ["a", "b", "c", "d", "e"]
    with pytest.raises(ValueError):
        merge_shard_results(5, [["a"], ["b"]])


def test_two_rank_gloo_sharding():
    n, world, port = 21, 2, 29517 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [hashlib.sha1(b).hexdigest() for b in _blobs(n)]
    for rank, keys, n_packed, tmax in results:
        assert keys == want                                   # every rank sees all keys, in input order
        assert n_packed == len(shard_indices(n, rank, world))   # but packed only its own shard
        assert tmax == 11.0                                   # max over ranks
