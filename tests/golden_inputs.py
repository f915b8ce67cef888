"""Deterministic inputs shared by oracle/make_golden.py (which runs the real reference
on them) and the tests (which re-create them and compare against tests/golden/)."""
import numpy as np


def artifact_objects():
    """name -> Python object, saved through TaskDataStore.save_artifacts."""
    rng = np.random.default_rng(1234)
    return {
        "none": None,
        "true": True,
        "answer": 42,
        "greeting": "hello metaflow",
        "config": {"alpha": 0.5, "layers": [64, 64, 10], "name": "mlp"},
        "words": ["foreach", "join", "step", "artifact"] * 64,
        "normal_64k": np.random.default_rng(0).standard_normal(16384, dtype=np.float32),
        "uniform_32k": rng.random(8192, dtype=np.float32),
        "lowent_64k": rng.integers(0, 16, 16384).astype(np.float32),
        "zeros_256k": np.zeros(65536, dtype=np.float32),
        "int64_ramp": np.arange(5000, dtype=np.int64),
        "dup_of_true": True,
    }


def raw_blobs():
    """name -> bytes, saved through ContentAddressedStore.save_blobs directly."""
    rng = np.random.default_rng(99)
    return {
        "empty": b"",
        "abc": b"abc",
        "one_block_minus": b"a" * 55,
        "one_block_edge": b"a" * 56,
        "two_blocks": b"a" * 64,
        "text": (b"The quick brown fox jumps over the lazy dog. " * 400),
        "random_5000": rng.bytes(5000),
        "ramp_70000": bytes(range(256)) * 273 + b"xyz",
    }


def large_artifacts():
    """name -> 16 MiB fp32 array (pickle blob 16,777,379 B) saved by the REAL reference in oracle/make_golden_large.py."""
    n = 4194304
    rng = np.random.default_rng(77)
    lowent = rng.integers(0, 16, n).astype(np.float32)
    sparse = np.zeros(n, dtype=np.float32)
    for k in range(0, n, 16384):                       # a 2 KiB stripe of noise at the start of every 64 KiB
        sparse[k:k + 512] = rng.standard_normal(512, dtype=np.float32)
    return {"lowent_16mib": lowent, "sparse_16mib": sparse}
