"""bench.py's contract that can be checked without a GPU: the reference arm prints ONE JSON line with the keys the driver
reads (it times oracle/ref_path.py on the host cores and touches no CUDA), under torchrun only rank 0 works, and the product
arm fails loudly on a machine without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from metaflow_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, capture_output=True, text=True, cwd=REPO, env=e, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--config", "C3", "--steps", "1", "--warmup", "1", "--ref-blobs-per-proc", "1"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "artifact_roundtrip_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert d["config"]["name"] == "C3" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank_zero_works():
    p = _run(["--impl", "reference", "--gpus", "2", "--config", "C3", "--steps", "1", "--warmup", "0", "--ref-blobs-per-proc", "1"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29599"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.skipif(capi.load().mfcas_device_count() > 0, reason="only meaningful without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    p = _run(["--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--api-blobs", "0"], timeout=300)
    assert p.returncode != 0
    assert "{\"metric\"" not in p.stdout
