"""Generate tests/golden/ref_written_16mib.json by running the REAL reference (imported from /root/reference)
on BASELINE-size artifacts (16 MiB fp32 arrays, pickle blob 16,777,379 B).

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_large.py
The files the reference writes for N(0,1) data are 15.6 MB each -- too big for a fixture -- so the arrays
here are chosen to keep the reference's gzip output small while still exercising what a stock zlib level-3
stream of that size contains (about a thousand dynamic-Huffman blocks, long and short matches, block
boundaries at arbitrary bit offsets):
  lowent : rng.integers(0, 16, n).astype(float32)            (SURVEY 8d secondary distribution, ratio 0.247)
  sparse : zeros with N(0,1) stripes every 64 KiB            (long zero runs + incompressible stretches)
For each: the recipe name, the key and the `info` the reference recorded, and the gzip FILE it wrote (base64).
N(0,1) artifacts of the same size are covered by the tests through oracle/ref_path.py (the same zlib call).
"""
import base64, json, os, shutil, sys, tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(REPO, "tests"))
os.environ.setdefault("USER", "golden")

import metaflow  # noqa: E402  (the reference)
from metaflow.datastore import FlowDataStore  # noqa: E402
from metaflow.plugins.datastores.local_storage import LocalStorage  # noqa: E402
import golden_inputs  # noqa: E402


def main():
    root = tempfile.mkdtemp(prefix="golden_large_")
    out = {"reference_version": metaflow.__version__, "artifacts": {}}
    try:
        fds = FlowDataStore("GoldenLarge", storage_impl=LocalStorage, ds_root=root)
        objs = golden_inputs.large_artifacts()
        w = fds.get_task_datastore("1", "start", "1", attempt=0, mode="w")
        w.init_task()
        w.save_artifacts(iter(objs.items()), len_hint=len(objs))
        w.done()
        with open(os.path.join(root, "GoldenLarge", "1", "start", "1", "0.data.json")) as f:
            dj = json.load(f)
        for n in objs:
            key = dj["objects"][n]
            with open(os.path.join(root, "GoldenLarge", "data", key[:2], key), "rb") as f:
                gz = f.read()
            out["artifacts"][n] = {"key": key, "info": dj["info"][n], "gz_len": len(gz), "gz_b64": base64.b64encode(gz).decode()}
            print(n, key, dj["info"][n]["size"], len(gz))
        back = dict(fds.get_task_datastore("1", "start", "1").load_artifacts(list(objs)))
        assert all((back[n] == objs[n]).all() for n in objs)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    dst = os.path.join(REPO, "tests", "golden", "ref_written_16mib.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
