"""Round-robin sharding of a step's artifact set over the GPUs / ranks of one box.

North-star sharding rule: blob i of the batch belongs to shard i mod G; blobs are
independent, so there is no collective on the data path -- the only exchange is the host-side gather of the
small per-blob results (keys, sizes), in input order because callers rely on it
(metaflow/datastore/task_datastore.py:382).  Inside one process libmfcas does this over the devices of its
context; across processes (one rank per GPU, as bench.py runs under torchrun) these helpers do.
"""


def shard_indices(n_items, rank, world):
    """Indices of the items shard `rank` of `world` owns: rank, rank + world, ..."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_items, world))


def merge_shard_results(n_items, per_rank_results):
    """per_rank_results[r] = results of shard r in shard order -> one list in input order."""
    world = len(per_rank_results)
    out = [None] * n_items
    for r, res in enumerate(per_rank_results):
        idx = shard_indices(n_items, r, world)
        if len(idx) != len(res):
            raise ValueError("shard %d returned %d results for %d items" % (r, len(res), len(idx)))
        for i, v in zip(idx, res):
            out[i] = v
    return out


def pack_sharded(codec, blobs, rank=0, world=1, all_gather=None):
    """Pack this rank's shard with `codec`; with all_gather (e.g. torch.distributed.all_gather_object bound to
    a list) the keys of all ranks are merged back to input order."""
    mine = shard_indices(len(blobs), rank, world)
    keys, packed = codec.pack_batch([blobs[i] for i in mine])
    if all_gather is None or world == 1:
        return keys, packed
    gathered = all_gather(keys)
    return merge_shard_results(len(blobs), gathered), packed
