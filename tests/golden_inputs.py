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
