"""Scatter pickling must be byte-identical to pickle.dumps(obj, protocol=4), the serialisation the reference
keys artifacts by (metaflow/datastore/task_datastore.py:351-355)."""
import pickle
from collections import OrderedDict, namedtuple

import numpy as np
import pytest

from metaflow_b200.serializer import Segments, dumps_segments

Point = namedtuple("Point", "x y")


class Holder(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _cases():
    rng = np.random.default_rng(7)
    big = rng.standard_normal(1 << 20, dtype=np.float32)           # 4 MiB
    edge = np.arange(16384, dtype=np.float32)                       # exactly 64 KiB: first size on the fast path
    below = np.arange(16383, dtype=np.float32)                      # just below it
    yield "big_f4", big
    yield "edge_64k", edge
    yield "below_64k", below
    yield "2d_f8", rng.standard_normal((300, 517))
    yield "3d_i2", rng.integers(-5, 5, size=(7, 64, 333), dtype=np.int16)
    yield "fortran", np.asfortranarray(rng.standard_normal((400, 300)))
    yield "non_contiguous", big[::3]
    yield "big_endian", big.astype(">f4")
    yield "bool", rng.integers(0, 2, size=200000).astype(bool)
    yield "complex", (rng.standard_normal(50000) + 1j * rng.standard_normal(50000))
    yield "datetime", np.arange(20000).astype("datetime64[s]")
    yield "structured", np.zeros(9000, dtype=[("a", "<i4"), ("b", "<f8"), ("c", "S3")])
    yield "unicode_dtype", np.array(["abc", "defgh"] * 9000)
    yield "object_dtype", np.array([{"a": 1}, None, "x"] * 30000, dtype=object)
    yield "zero_size", np.empty((0, 5), dtype=np.float64)
    yield "scalar_array", np.float32(3.5) * np.ones(())
    yield "same_array_twice", [big, big, {"again": big}]
    yield "two_arrays_same_dtype", (big, big.copy(), edge)
    yield "view_and_base", [big, big[: 1 << 19]]
    yield "dict_mixed", {"w": big, "b": below, "n": 3, "s": "text", "l": list(range(1000)), "t": (1, 2.5, None)}
    yield "holder", Holder(weights=big, meta={"epoch": 3}, p=Point(1, 2))
    yield "ordered", OrderedDict([("a", edge), ("b", big)])
    yield "large_bytes", b"\x01" * 300000
    yield "large_bytes_zero_64k", bytes(1 << 16)                     # looks like the placeholder, is not it
    yield "bytes_and_array", [bytes(1 << 16), edge, bytes(1 << 16)]
    yield "bytearray", bytearray(b"ab" * 100000)
    yield "string_large", "x" * 200000
    yield "plain", {"a": [1, 2, 3], "b": None}
    yield "matrix_subclass", np.ma.masked_array(rng.standard_normal(40000), mask=False)
    yield "many_small", [np.arange(10) for _ in range(100)]
    yield "nested_lists", [[edge, [big]], (edge,)]


@pytest.mark.parametrize("name,obj", list(_cases()), ids=[n for n, _ in _cases()])
def test_identical_to_pickle_dumps(name, obj):
    want = pickle.dumps(obj, protocol=4)
    seg = dumps_segments(obj)
    assert isinstance(seg, Segments)
    assert len(seg) == len(want)
    assert seg.tobytes() == want
    dst = np.full(len(want) + 64, 0xEE, dtype=np.uint8)
    seg.copy_into(dst, 32)
    assert dst[32:32 + len(want)].tobytes() == want
    assert (dst[:32] == 0xEE).all() and (dst[32 + len(want):] == 0xEE).all()


def test_large_arrays_are_views_not_copies():
    a = np.arange(1 << 20, dtype=np.float32)
    seg = dumps_segments({"a": a})
    views = [p for p in seg.parts if isinstance(p, np.ndarray)]
    assert len(views) == 1 and np.shares_memory(views[0], a)
    assert sum(len(p) for p in seg.parts if not isinstance(p, np.ndarray)) < 1024


def test_pandas_frame_if_available():
    pd = pytest.importorskip("pandas")
    df = pd.DataFrame({"a": np.arange(200000, dtype=np.float64), "b": np.arange(200000, dtype=np.int64), "c": ["s"] * 200000})
    assert dumps_segments(df).tobytes() == pickle.dumps(df, protocol=4)


def test_torch_tensor_if_available():
    torch = pytest.importorskip("torch")
    t = torch.arange(100000, dtype=torch.float32)
    assert pickle.loads(dumps_segments(t).tobytes()).equal(t)
    assert len(dumps_segments(t)) == len(pickle.dumps(t, protocol=4))


def test_unpicklable_raises_like_pickle():
    with pytest.raises((TypeError, AttributeError, pickle.PicklingError)):
        dumps_segments({"f": (lambda: 0)})
    import threading
    with pytest.raises(TypeError):
        dumps_segments(threading.Lock())


# ---- gather unpickling
from concurrent.futures import ThreadPoolExecutor

from metaflow_b200.serializer import loads_batch


def _same(a, b):
    if isinstance(a, np.ndarray):
        return type(a) is type(b) and a.dtype == b.dtype and a.shape == b.shape and (a.dtype.hasobject or a.tobytes() == b.tobytes())
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, dict):
        return type(a) is type(b) and list(a) == list(b) and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, Holder):
        return _same(a.__dict__, b.__dict__)
    return type(a) is type(b) and (a == b if not hasattr(a, "__array__") else np.array_equal(a, b))


def test_loads_batch_equals_pickle_loads():
    names, objs = zip(*_cases())
    blobs = [pickle.dumps(o, protocol=4) for o in objs]
    views = [memoryview(np.frombuffer(b, dtype=np.uint8).copy()) for b in blobs]     # like the pinned arena: not bytes
    with ThreadPoolExecutor(4) as pool:
        for got in (loads_batch(views, pool), loads_batch(views, None), loads_batch(blobs, None)):
            for name, g, v in zip(names, got, views):
                want = pickle.loads(v)
                assert _same(want, g), name


def test_loads_batch_keeps_sharing_and_writability():
    a = np.arange(1 << 16, dtype=np.float64)
    blob = pickle.dumps([a, a, {"k": a}], protocol=4)
    out = loads_batch([memoryview(blob)])[0]
    assert out[0] is out[1] is out[2]["k"]
    assert out[0].flags.writeable == pickle.loads(blob)[0].flags.writeable
    out[0][0] = 5.0                        # a fresh object: writing to it must not touch the source stream
    assert pickle.loads(blob)[0][0] == 0.0
    # the same blob twice gives two independent objects
    x, y = loads_batch([memoryview(blob), memoryview(blob)])
    x[0][1] = 9.0
    assert y[0][1] == 1.0


def test_loads_batch_other_protocols_and_garbage():
    a = np.arange(50000, dtype=np.float32)
    for proto in (2, 3, 5):
        assert np.array_equal(loads_batch([pickle.dumps(a, protocol=proto)])[0], a)
    with pytest.raises(Exception):
        loads_batch([b"\x80\x04\x95\xff\xff\xff\xff\xff\xff\xff\x7fgarbage"])
    with pytest.raises(Exception):
        loads_batch([b"not a pickle"])
