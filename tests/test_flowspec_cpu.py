"""BASELINE config C1 (plumbing, no GPU): a FlowSpec with an 8-way foreach of 1 MiB numpy artifacts and a
join, run by the in-process runner with the test-only CPU codec; the datastore it leaves behind is read
back, and -- in the build container -- compared with what the REAL reference writes for the same flow."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle_codec import OracleCodec
from metaflow_b200.flowspec import FlowRunner, FlowSpec, MergeArtifactsException, step

HAVE_REF = os.path.isdir("/root/reference/metaflow")

FLOW_SRC = textwrap.dedent('''
    import numpy as np
    from {pkg} import FlowSpec, step

    class ForeachFlow(FlowSpec):
        @step
        def start(self):
            self.seeds = list(range(8))
            self.common = "shared-by-all"
            self.next(self.work, foreach="seeds")

        @step
        def work(self):
            self.arr = np.random.default_rng(self.input).standard_normal(262144, dtype=np.float32)   # 1 MiB
            self.seed = self.input
            self.next(self.join)

        @step
        def join(self, inputs):
            self.total = float(sum(float(inp.arr.sum()) for inp in inputs))
            self.order = [inp.seed for inp in inputs]
            self.merge_artifacts(inputs, exclude=["arr", "seed"])
            self.next(self.end)

        @step
        def end(self):
            self.done_flag = True

    if __name__ == "__main__":
        ForeachFlow()
''')


def _load_flow(tmp_path):
    src = FLOW_SRC.format(pkg="metaflow_b200.flowspec")
    ns = {}
    exec(compile(src, str(tmp_path / "flow_b200.py"), "exec"), ns)
    return ns["ForeachFlow"]


def test_foreach_join_flow_on_cpu(tmp_path):
    codec = OracleCodec()
    run = FlowRunner(_load_flow(tmp_path), ds_root=str(tmp_path / "ds"), codec=codec).run()
    assert codec.calls["pack"] == 4                              # one batch per STEP (start, work x8, join, end)
    work = run.task_datastores("work")
    assert len(work) == 8
    want = [np.random.default_rng(i).standard_normal(262144, dtype=np.float32) for i in range(8)]
    for i, ds in enumerate(work):
        assert ds["seed"] == i and np.array_equal(ds["arr"], want[i])
        assert ds["_foreach_stack"][0].index == i and ds["_foreach_stack"][0].num_splits == 8
        assert ds.pathspec_index == "1/work[%d]" % i
        assert ds["common"] == "shared-by-all"
    j = run.task_datastores("join")[0]
    assert j["order"] == list(range(8))
    assert abs(j["total"] - float(sum(float(a.sum()) for a in want))) < 1e-3
    assert j["common"] == "shared-by-all" and "arr" not in j
    assert run.task_datastores("end")[0]["done_flag"] is True
    # the datastore layout is Metaflow's
    tdir = tmp_path / "ds" / "ForeachFlow" / "1" / "work" / work[3].task_id
    dj = json.loads((tdir / "0.data.json").read_text())
    assert dj["version"] == "1.0" and set(("arr", "seed", "_foreach_stack", "_transition", "name")) <= set(dj["objects"])
    assert (tdir / "0.DONE.lock").read_text() == '""'


def test_static_split_and_merge_conflict(tmp_path):
    class SplitFlow(FlowSpec):
        @step
        def start(self):
            self.x = 1
            self.next(self.a, self.b)

        @step
        def a(self):
            self.y = "from-a"
            self.next(self.join)

        @step
        def b(self):
            self.y = "from-b"
            self.next(self.join)

        @step
        def join(self, inputs):
            self.ys = sorted(i.y for i in inputs)
            self.via_name = inputs.a.y
            self.merge_artifacts(inputs, exclude=["y"])
            self.next(self.end)

        @step
        def end(self):
            pass

    run = FlowRunner(SplitFlow, ds_root=str(tmp_path / "ds"), codec=OracleCodec()).run()
    j = run.task_datastores("join")[0]
    assert j["ys"] == ["from-a", "from-b"] and j["via_name"] == "from-a" and j["x"] == 1

    class BadMerge(SplitFlow):
        @step
        def join(self, inputs):
            self.merge_artifacts(inputs)
            self.next(self.end)

    with pytest.raises(MergeArtifactsException):
        FlowRunner(BadMerge, ds_root=str(tmp_path / "ds2"), codec=OracleCodec()).run()


def test_nested_foreach(tmp_path):
    class Nested(FlowSpec):
        @step
        def start(self):
            self.outer = ["a", "b", "c"]
            self.next(self.mid, foreach="outer")

        @step
        def mid(self):
            self.inner = [1, 2]
            self.next(self.leaf, foreach="inner")

        @step
        def leaf(self):
            self.tag = "%s%d" % (self.foreach_stack()[0][2], self.input)
            self.next(self.inner_join)

        @step
        def inner_join(self, inputs):
            self.tags = [i.tag for i in inputs]
            self.next(self.outer_join)

        @step
        def outer_join(self, inputs):
            self.all_tags = [t for i in inputs for t in i.tags]
            self.next(self.end)

        @step
        def end(self):
            pass

    run = FlowRunner(Nested, ds_root=str(tmp_path / "ds"), codec=OracleCodec()).run()
    assert len(run.task_datastores("leaf")) == 6 and len(run.task_datastores("inner_join")) == 3
    assert run.task_datastores("outer_join")[0]["all_tags"] == ["a1", "a2", "b1", "b2", "c1", "c2"]


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")
def test_same_artifact_keys_as_the_real_reference(tmp_path):
    """The reference runs the same flow file through its own CLI + scheduler (process per task); user
    artifacts must get identical content-address keys, and the reference's client-side loader must read ours."""
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    (ref_dir / "flow.py").write_text(FLOW_SRC.format(pkg="metaflow"))
    env = dict(os.environ, PYTHONPATH="/root/reference", PYTHONDONTWRITEBYTECODE="1", USER="c1test",
               METAFLOW_DATASTORE_SYSROOT_LOCAL=str(ref_dir), METAFLOW_DEFAULT_METADATA="local")
    out = subprocess.run([sys.executable, "flow.py", "--no-pylint", "--datastore=local", "--metadata=local", "run"],
                         cwd=str(ref_dir), env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    root = ref_dir / ".metaflow" / "ForeachFlow"
    run_id = sorted(p.name for p in root.iterdir() if p.name.isdigit())[-1]
    ref_keys = {}
    for tdir in (root / run_id / "work").iterdir():
        if not (tdir / "0.data.json").exists():
            continue                      # the _meta directory of the local metadata provider
        dj = json.loads((tdir / "0.data.json").read_text())
        ref_keys[dj["objects"]["seed"]] = dj["objects"]
    ours = FlowRunner(_load_flow(tmp_path), ds_root=str(tmp_path / "ds"), codec=OracleCodec()).run()
    for ds in ours.task_datastores("work"):
        mine = ds._objects
        theirs = ref_keys[mine["seed"]]
        for name in ("arr", "seed", "common", "seeds"):
            assert mine[name] == theirs[name], name


def test_wide_foreach_is_still_one_batch_per_step(tmp_path):
    """The reference's `wide_foreach` core test shape (1200 splits): every sibling's artifacts go to the codec as one
    batch per step, the join sees all inputs in order, and identical artifacts of the siblings are stored once."""
    from metaflow_b200.flowspec import FlowSpec, step

    class WideFlow(FlowSpec):
        @step
        def start(self):
            self.items = list(range(1200))
            self.next(self.each, foreach="items")

        @step
        def each(self):
            self.sq = self.input * self.input
            self.same = "identical in every split"
            self.next(self.join)

        @step
        def join(self, inputs):
            self.n = len(inputs)
            self.total = sum(inp.sq for inp in inputs)
            self.first_last = (inputs[0].sq, inputs[-1].sq)
            self.next(self.end)

        @step
        def end(self):
            pass

    codec = OracleCodec()
    run = FlowRunner(WideFlow, ds_root=str(tmp_path / "ds"), codec=codec).run()
    assert codec.calls["pack"] == 4                                  # start, each x1200, join, end
    j = run.task_datastores("join")[0]
    assert j["n"] == 1200 and j["total"] == sum(i * i for i in range(1200)) and j["first_last"] == (0, 1199 * 1199)
    each = run.task_datastores("each")
    assert len(each) == 1200 and each[777]["sq"] == 777 * 777
    keys = {ds._objects["same"] for ds in each}
    assert len(keys) == 1                                            # content-addressed: one blob for 1200 identical artifacts
    blob_files = [f for _, _, fs in os.walk(tmp_path / "ds" / "WideFlow" / "data") for f in fs if not f.endswith("_meta")]
    assert len(blob_files) == len(set(blob_files))


def test_top_level_api_surface():
    """`from metaflow_b200 import FlowSpec, step` -- the names a flow file imports from `metaflow`."""
    import metaflow_b200 as mf
    for name in ("FlowSpec", "step", "FlowRunner", "FlowDataStore", "TaskDataStore", "ContentAddressedStore", "LocalStorage",
                 "TaskDataStoreSet", "Inputs", "DataException", "FileCache"):
        assert getattr(mf, name) is not None, name
    with pytest.raises(AttributeError):
        mf.no_such_name
