"""Generate tests/golden/ by running the REAL reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
Writes tests/golden/cas_golden.json: for every input in tests/golden_inputs.py the key the
reference computed, the gzip file the reference wrote (base64), its _meta sidecar, and
for the TaskDataStore leg the reference's data.json objects/info maps.
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
    root = tempfile.mkdtemp(prefix="golden_ds_")
    out = {"reference_version": metaflow.__version__, "raw_blobs": {}, "artifacts": {}}
    try:
        fds = FlowDataStore("GoldenFlow", storage_impl=LocalStorage, ds_root=root)
        # --- CAS level: save_blobs(raw=False) on raw byte strings
        blobs = golden_inputs.raw_blobs()
        names = list(blobs)
        res = fds.ca_store.save_blobs(iter(blobs[n] for n in names), len_hint=len(names))
        for n, r in zip(names, res):
            path = os.path.join(root, "GoldenFlow", "data", r.key[:2], r.key)
            with open(path, "rb") as f:
                gz = f.read()
            with open(path + "_meta") as f:
                meta = json.load(f)
            assert r.uri is None
            out["raw_blobs"][n] = {"key": r.key, "len": len(blobs[n]), "meta": meta,
                                   "gz_b64": base64.b64encode(gz).decode()}
        # round trip through the reference's own load_blobs
        got = dict(fds.ca_store.load_blobs([out["raw_blobs"][n]["key"] for n in names]))
        for n in names:
            assert got[out["raw_blobs"][n]["key"]] == blobs[n]
        # --- TaskDataStore level
        objs = golden_inputs.artifact_objects()
        w = fds.get_task_datastore("1", "start", "7", attempt=0, mode="w")
        w.init_task()
        w.save_artifacts(iter(objs.items()), len_hint=len(objs))
        w.done()
        with open(os.path.join(root, "GoldenFlow", "1", "start", "7", "0.data.json")) as f:
            data_json = json.load(f)
        out["data_json"] = {k: data_json[k] for k in ("datastore", "version", "attempt", "objects", "info")}
        with open(os.path.join(root, "GoldenFlow", "1", "start", "7", "0.DONE.lock"), "rb") as f:
            out["done_lock_b64"] = base64.b64encode(f.read()).decode()
        for n in objs:
            key = data_json["objects"][n]
            path = os.path.join(root, "GoldenFlow", "data", key[:2], key)
            with open(path, "rb") as f:
                gz = f.read()
            out["artifacts"][n] = {"key": key, "gz_b64": base64.b64encode(gz).decode()}
        r = fds.get_task_datastore("1", "start", "7")
        back = dict(r.load_artifacts(list(objs)))
        assert set(back) == set(objs)
        # error-message shapes pinned by test/unit/test_content_addressed_store.py
        out["error_shapes"] = {
            "missing_version": "Could not extract encoding version for '%s'",
            "unknown_version": "Unknown encoding version %d for '%s' -- the artifact is either corrupt or you need to update Metaflow to the latest version",
            "unpack_failure": "Could not unpack artifact '%s': %s",
        }
    finally:
        shutil.rmtree(root, ignore_errors=True)
    dst = os.path.join(REPO, "tests", "golden", "cas_golden.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
