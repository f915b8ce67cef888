#!/usr/bin/env python
"""bench.py -- artifact GB/s through save+load of the content-addressed store (BASELINE.json metric).

Workload (BASELINE.json configs[1]): a 4096-way foreach whose tasks each persist one 16 MiB fp32
N(0,1) numpy array (pickle protocol 4 -> 16,777,379-byte blob).  One STEP = one wave of
`--blobs-per-step` such blobs per GPU going through the whole hot path:
    save: SHA-1 key + DEFLATE + CRC-32 -> RFC 1952 file      (mfcas_pack_*)
    load: INFLATE + CRC-32/ISIZE verification -> blob bytes  (mfcas_unpack_*)
With the defaults (1024 blobs/step, 4 steps) the job is exactly the 4096 blobs (SHA-1 is one serial chain
per blob, so the wave is made as wide as HBM allows: 1024 concurrent chains).  Round-robin sharding:
blob i of the global batch belongs to rank i mod N (weak scaling: per-GPU wave size is fixed).

  value : round-trip GB/s with inputs/outputs resident in HBM (mfcas_pack_device_extents / _unpack_device),
          CUDA events on the library stream, max over ranks.
  e2e   : the same metric through the host-pointer C ABI (mfcas_pack_batch / _unpack_batch) from
          page-locked host buffers, H2D and D2H inside the timed region.
  --impl reference : the reference's CPU path (hashlib.sha1 + gzip level 3 + gunzip + local CAS
          files, restated in oracle/ref_path.py) on all host cores, bounded sample per step.
"""
import argparse
import ctypes
import json
import os
import pickle
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# one hardware work queue per in-flight wave stream (must be set before the CUDA context is created)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

ARRAY_ELEMS = 4194304            # 16 MiB of float32 (config C2; --config C3 / C4 change it)
METRIC = "artifact_roundtrip_GBps"
UNIT = "GB/s"

# BASELINE.json configs that fit one GPU: elements per fp32 artifact, blobs per step per GPU, what it is
CONFIGS = {
    # (3072 per step: SHA-1 is one serial chain per blob -- 116 ms for 16 MiB however wide the wave is -- so the step is
    # made as wide as the 180 GB of HBM allow: input + packed slots + decoded output of 3072 blobs = 155 GB)
    "C2": (4194304, 3072, "C2 foreach x4096, 16 MiB fp32 N(0,1) artifacts"),
    "C3": (16384, 65536, "C3 foreach x65536, 64 KiB fp32 N(0,1) artifacts (hash-bound small-blob path)"),
    "C4": (262144, 16384, "C4 nested foreach 256x256, 1 MiB fp32 N(0,1) artifacts, join loads (65536 leaves = 4 steps)"),
}


def pickle_frame(n_elems):
    """(prefix, suffix, nbytes): pickle.dumps(float32[n], protocol=4) == prefix + raw + suffix."""
    probe = np.full(n_elems, 1.5, dtype=np.float32)
    p = pickle.dumps(probe, protocol=4)
    raw = probe.tobytes()
    off = p.index(raw[:64])
    assert p[off:off + len(raw)] == raw
    return p[:off], p[off + len(raw):], len(raw)


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(object):
    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def usable_cores():
    """Host threads this process may really use: cpu_count, affinity mask and the cgroup CPU quota."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as f:
                parts = f.read().split()
            if path.endswith("cpu.max"):
                if parts[0] != "max":
                    n = min(n, max(1, int(int(parts[0]) / int(parts[1]) + 0.5)))
            else:
                q = int(parts[0])
                if q > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f2:
                        n = min(n, max(1, int(q / int(f2.read().split()[0]) + 0.5)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


# ----------------------------------------------------------------------------- reference (CPU) arm
def _cpu_worker(args):
    """One process = one reference task: generate, save through the restated path, load back."""
    worker, n_blobs, root, seed0 = args
    from oracle import ref_path
    blobs = [pickle.dumps(np.random.default_rng(seed0 + worker * 1000 + j).standard_normal(ARRAY_ELEMS, dtype=np.float32), protocol=4)
             for j in range(n_blobs)]
    t0 = time.perf_counter()
    res = ref_path.save_blobs(root, "BenchFlow/data", blobs)
    t1 = time.perf_counter()
    back = ref_path.load_blobs(root, "BenchFlow/data", [k for _, k in res])
    t2 = time.perf_counter()
    ok = all(b == o for b, (_, o) in zip(blobs, back))
    return sum(map(len, blobs)), t1 - t0, t2 - t1, ok


def cpu_reference_pass(n_procs, blobs_per_proc, seed0):
    """Process-per-task parallelism of the reference (runtime.py:1517-1540, --max-workers) on a tmpfs
    datastore.  Returns (bytes, t_save, t_load): slowest worker's in-call times."""
    import multiprocessing as mp
    import shutil
    import tempfile
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    root = tempfile.mkdtemp(prefix="mfcas_ref_", dir=base)
    try:
        ctx = mp.get_context("fork")
        with ctx.Pool(n_procs) as pool:
            out = pool.map(_cpu_worker, [(w, blobs_per_proc, root, seed0) for w in range(n_procs)])
    finally:
        shutil.rmtree(root, ignore_errors=True)
    assert all(o[3] for o in out), "reference arm round-trip mismatch"
    return sum(o[0] for o in out), max(o[1] for o in out), max(o[2] for o in out)


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = usable_cores()
    per = max(1, args.ref_blobs_per_proc)
    for i in range(args.warmup):
        cpu_reference_pass(cores, 1 if i else 1, 7_000_000 + i)
    tot_b = tot_s = tot_l = 0.0
    for s in range(args.steps):
        b, ts, tl = cpu_reference_pass(cores, per, 1_000_000 * (s + 1))
        tot_b += b; tot_s += ts; tot_l += tl
    value = tot_b / (tot_s + tot_l) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (tot_s + tot_l) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "impl": "reference",
        "config": {"workload": "%s, save_blobs+load_blobs" % args.workload, "name": args.config, "sample": "%d blobs/step (%d procs x %d)" % (cores * per, cores, per),
                   "datastore": "local on tmpfs"},
        "save_GBps": tot_b / tot_s / 1e9, "load_GBps": tot_b / tot_l / 1e9,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "logical_cpus": os.cpu_count(), "kind": "port",
                         "sample": "%d blobs per step through oracle/ref_path.py (hashlib.sha1 + gzip level 3 + gunzip + tmpfs CAS files), process per task" % (cores * per)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    from metaflow_b200 import capi
    lib = capi.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    devs = (ctypes.c_int * 1)(local_rank)
    ctx = lib.mfcas_create(devs, 1, args.wave_mib << 20, 0)
    if not ctx:
        raise SystemExit("mfcas_create failed: no CUDA device -- the hot path has no CPU fallback")
    ctx = ctypes.c_void_p(ctx)
    prefix, suffix, nbytes = pickle_frame(ARRAY_ELEMS)
    blob_len = len(prefix) + nbytes + len(suffix)
    B = args.blobs_per_step
    stride = (blob_len + 255) // 256 * 256 + 256
    cap = lib.mfcas_pack_bound(blob_len)
    cstride = (cap + 255) // 256 * 256 + 256

    # HBM-resident leg = the product's own save call in its device form: every 2 MiB stripe of a blob is its own gzip
    # member, written by the encoder straight into a fixed slot of the blob's destination (mfcas_pack_device_extents;
    # GpuCodec.pack_batch makes the same call on host buffers and a writer writev()s the extents).  Between the timed
    # save and the timed load the extents are moved together in place (mfcas_compact_device) -- the file write of the
    # real path, not timed here, exactly as in the e2e leg below.
    ecount_d = lib.mfcas_pack_extent_count(blob_len)
    cap = lib.mfcas_pack_extent_bound(blob_len)
    cstride = (cap + 255) // 256 * 256 + 256
    d_in = torch.empty(B * stride, dtype=torch.uint8, device=dev)
    d_packed = torch.empty(B * cstride, dtype=torch.uint8, device=dev)
    d_out = torch.empty(B * stride, dtype=torch.uint8, device=dev)
    d_rand = torch.empty(ARRAY_ELEMS, dtype=torch.float32, device=dev)
    t_prefix = torch.frombuffer(bytearray(prefix), dtype=torch.uint8).to(dev)
    t_suffix = torch.frombuffer(bytearray(suffix), dtype=torch.uint8).to(dev)
    gen = torch.Generator(device=dev)

    def generate(step):
        # blob i of this rank's wave is global blob (step*B + i)*world + rank  (round-robin i mod N)
        for i in range(B):
            gen.manual_seed(((step * B + i) * world + rank) * 2654435761 % (1 << 62))
            d_rand.normal_(generator=gen)
            o = i * stride
            d_in[o:o + len(prefix)] = t_prefix
            d_in[o + len(prefix):o + len(prefix) + nbytes] = d_rand.view(torch.uint8)
            d_in[o + len(prefix) + nbytes:o + blob_len] = t_suffix
        torch.cuda.synchronize()

    u64, vp = ctypes.c_uint64, ctypes.c_void_p
    n = B
    src = (vp * n)(*[d_in.data_ptr() + i * stride for i in range(n)])
    slen = (u64 * n)(*([blob_len] * n))
    pk = (vp * n)(*[d_packed.data_ptr() + i * cstride for i in range(n)])
    pcap = (u64 * n)(*([cap] * n))
    plen = (u64 * n)()
    d_ext_off = (u64 * (n * ecount_d))()
    d_ext_len = (u64 * (n * ecount_d))()
    out = (vp * n)(*[d_out.data_ptr() + i * stride for i in range(n)])
    ocap = (u64 * n)(*([blob_len] * n))
    olen = (u64 * n)()
    st = (ctypes.c_int32 * n)()
    sha = ctypes.create_string_buffer(20 * n)
    stream = torch.cuda.ExternalStream(lib.mfcas_stream(ctx, 0), device=dev)

    def check(rc, what):
        if rc != 0 or any(st):
            raise SystemExit("%s failed rc=%d status=%s: %s" % (what, rc, [s for s in st if s][:4], lib.mfcas_last_error(ctx)))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        e0.record(stream)
        check(lib.mfcas_pack_device_extents(ctx, 0, n, src, slen, pk, pcap, d_ext_off, d_ext_len, sha, st), "pack_device_extents")
        e1.record(stream)
        ms = (ctypes.c_float * 8)()
        lib.mfcas_last_kernel_ms(ctx, ms)
        ms_pack = list(ms)
        rc = lib.mfcas_compact_device(ctx, 0, n, pk, slen, d_ext_len, plen)             # untimed: "the file write"
        if rc != 0:
            raise SystemExit("compact_device failed rc=%d: %s" % (rc, lib.mfcas_last_error(ctx)))
        e2.record(stream)
        check(lib.mfcas_unpack_device(ctx, 0, n, pk, plen, out, ocap, olen, st), "unpack_device")
        e3.record(stream)
        torch.cuda.synchronize()
        lib.mfcas_last_kernel_ms(ctx, ms)
        ms_all = [a if k < 3 else b for k, (a, b) in enumerate(zip(ms_pack, list(ms)))]
        return e0.elapsed_time(e1), e2.elapsed_time(e3), ms_all

    def verify(sample):
        import gzip
        import hashlib
        for i in sample:
            blob = d_in[i * stride:i * stride + blob_len].cpu().numpy().tobytes()
            assert sha.raw[20 * i:20 * i + 20] == hashlib.sha1(blob).digest(), "sha1 mismatch blob %d" % i
            z = d_packed[i * cstride:i * cstride + plen[i]].cpu().numpy().tobytes()
            assert gzip.decompress(z) == blob, "reference reader cannot decode blob %d" % i
            assert d_out[i * stride:i * stride + blob_len].cpu().numpy().tobytes() == blob, "decoded bytes differ blob %d" % i
            arr = pickle.loads(blob)
            assert arr.dtype == np.float32 and arr.shape == (ARRAY_ELEMS,)

    # ---- warm-up (untimed) + parity gate
    for w in range(max(1, args.warmup)):      # at least one untimed pass: it is also the parity gate
        generate(10_000 + w)
        step_device()
    # parity gate: every blob for the small-blob config, >= 1 % of the step otherwise -- keys against
    # hashlib, our files through the reference's reader (CPython gzip), decoded bytes against the inputs
    n_check = n if args.config == "C3" else max(3, int(n * args.verify_frac + 0.999))
    picks = np.random.default_rng(1234 + rank).choice(n, size=min(n, n_check), replace=False)
    verify(sorted(set(int(i) for i in picks) | {0, n - 1}))
    verified = len(set(int(i) for i in picks) | {0, n - 1})

    # ---- HBM-resident timed steps
    counters0 = [u64(0), u64(0), u64(0)]
    lib.mfcas_counters(ctx, *[ctypes.byref(c) for c in counters0])
    sampler = ClockSampler(local_rank)
    t_save = t_load = 0.0
    kms = [0.0] * 8
    packed_bytes = 0
    barrier()
    sampler.start()
    for s in range(args.steps):
        generate(s)
        a, b, ms = step_device()
        t_save += a; t_load += b
        kms = [x + y for x, y in zip(kms, ms)]
        packed_bytes += sum(plen)
    barrier()
    clocks = sampler.stop()
    counters1 = [u64(0), u64(0), u64(0)]
    lib.mfcas_counters(ctx, *[ctypes.byref(c) for c in counters1])
    launches = counters1[0].value - counters0[0].value
    picks = np.random.default_rng(4321 + rank).choice(n, size=min(n, max(3, int(n * args.verify_frac + 0.999))), replace=False)
    verify(sorted(int(i) for i in picks))               # and again on the last timed step's output
    verified += len(picks)
    raw_bytes = args.steps * n * blob_len

    # ---- end-to-end through the host-pointer C ABI (pinned host buffers, copies timed).
    # Save = mfcas_pack_batch_extents (the call GpuCodec.pack_batch makes: 2 MiB stripes, every stripe its own
    # gzip member in a fixed slot; a writer writev()s the extents into the CAS file).  Between the timed save and
    # the timed load the extents are made contiguous on the host -- that is the file write of the real path and is
    # not timed here -- and load = mfcas_unpack_batch on whole packed blobs.
    ecount = lib.mfcas_pack_extent_count(blob_len)
    ecap = lib.mfcas_pack_extent_bound(blob_len)
    estride = (ecap + 255) // 256 * 256 + 256
    # the step's blobs go through the host API in sub-batches of <= 1024 (17 GB): the path is PCIe-bound, so its
    # rate does not depend on the batch, and the page-locked buffers stay at 3 x 17 GB per rank
    ne = n
    while ne * stride > (18 << 30) and ne % 2 == 0:
        ne //= 2
    assert n % ne == 0
    h_in_p = lib.mfcas_alloc_pinned(ne * stride); h_pk_p = lib.mfcas_alloc_pinned(ne * estride); h_out_p = lib.mfcas_alloc_pinned(ne * stride)
    assert h_in_p and h_pk_p and h_out_p, "pinned allocation failed"
    h_in = np.ctypeslib.as_array((ctypes.c_uint8 * (ne * stride)).from_address(h_in_p))
    h_out = np.ctypeslib.as_array((ctypes.c_uint8 * (ne * stride)).from_address(h_out_p))
    h_pk = np.ctypeslib.as_array((ctypes.c_uint8 * (ne * estride)).from_address(h_pk_p))
    hsrc = (vp * ne)(*[h_in_p + i * stride for i in range(ne)])
    hpk = (vp * ne)(*[h_pk_p + i * estride for i in range(ne)])
    hout = (vp * ne)(*[h_out_p + i * stride for i in range(ne)])
    epcap = (u64 * ne)(*([ecap] * ne))
    eslen = (u64 * ne)(*([blob_len] * ne))
    eplen = (u64 * ne)()
    eocap = (u64 * ne)(*([blob_len] * ne))
    eolen = (u64 * ne)()
    est = (ctypes.c_int32 * ne)()
    esha = ctypes.create_string_buffer(20 * ne)
    ext_off = (u64 * (ne * ecount))()
    ext_len = (u64 * (ne * ecount))()
    t_in = torch.from_numpy(h_in)
    e2e_save = e2e_load = 0.0
    e2e_steps = args.steps
    h2d = d2h = 0

    def echeck(rc, what):
        if rc != 0 or any(est):
            raise SystemExit("%s failed rc=%d status=%s: %s" % (what, rc, [x for x in est if x][:4], lib.mfcas_last_error(ctx)))

    def save_host():
        t0 = time.perf_counter()
        echeck(lib.mfcas_pack_batch_extents(ctx, ne, hsrc, eslen, hpk, epcap, ext_off, ext_len, esha, est), "pack_batch_extents")
        return time.perf_counter() - t0

    hpk_load = (vp * ne)()

    def compact_host():
        """extents -> one contiguous packed blob each, laid out back to back the way GpuCodec stages files for a load (256-byte
        aligned slots with a 256-byte gap: the layout the library's device arena uses too, so a wave uploads as one copy).
        This is what writev into the CAS file + reading the files back does on the real path; not timed."""
        eo = np.frombuffer(ext_off, dtype=np.uint64).reshape(ne, ecount)
        el = np.frombuffer(ext_len, dtype=np.uint64).reshape(ne, ecount)
        out = 0
        for i in range(ne):
            base, cur = i * estride, out
            for k in range(ecount):
                o, ln = int(eo[i, k]), int(el[i, k])
                if base + o != cur:
                    h_pk[cur:cur + ln] = h_pk[base + o:base + o + ln]        # always moves towards lower addresses
                cur += ln
            eplen[i] = cur - out
            hpk_load[i] = h_pk_p + out
            out = (cur + 255) // 256 * 256 + 256

    def load_host():
        t0 = time.perf_counter()
        echeck(lib.mfcas_unpack_batch(ctx, ne, hpk_load, eplen, hout, eocap, eolen, est), "unpack_batch")
        return time.perf_counter() - t0

    def refresh(part):
        t_in.copy_(d_in[part * ne * stride:(part + 1) * ne * stride])     # device -> page-locked host buffer (untimed)

    # ---- copy-only ceiling of this box for this rank COUNT: the same page-locked buffers, both directions at once (a save
    # uploads N while it downloads C, a load the reverse), all ranks at the same time, no kernels.  The e2e number is read
    # against it: what the copies allow is what the host's memory system and PCIe allow, whatever the kernels do.
    def copy_probe():
        x = min(ne * stride, 4 << 30)
        t_out_h = torch.from_numpy(h_out)
        sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        best = 0.0
        for _ in range(3):
            barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            with torch.cuda.stream(sa):
                ev[0].record(sa); d_in[:x].copy_(t_in[:x], non_blocking=True); ev[1].record(sa)
            with torch.cuda.stream(sb):
                ev[2].record(sb); t_out_h[:x].copy_(d_out[:x], non_blocking=True); ev[3].record(sb)
            torch.cuda.synchronize()
            ms = max(ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]))
            best = max(best, x / (ms * 1e-3) / 1e9)
        return best                                        # GB/s per direction with both directions busy

    both_each = copy_probe()
    generate(20_000); refresh(0); torch.cuda.synchronize(); save_host(); compact_host(); load_host()       # warm-up (arena allocation)
    import gzip
    import hashlib
    blob0 = h_in[:blob_len].tobytes()
    assert esha.raw[:20] == hashlib.sha1(blob0).digest() and h_out[:blob_len].tobytes() == blob0, "e2e parity gate failed"
    assert gzip.decompress(h_pk[:eplen[0]].tobytes()) == blob0, "reference reader cannot decode the striped blob"   # (blob 0 starts at offset 0)
    barrier()
    ec0 = [u64(0), u64(0), u64(0)]
    lib.mfcas_counters(ctx, *[ctypes.byref(c) for c in ec0])
    for s in range(e2e_steps):
        generate(30_000 + s)
        for part in range(n // ne):
            refresh(part)
            barrier()
            a = save_host()
            compact_host()
            barrier()
            b = load_host()
            e2e_save += a; e2e_load += b
    ec1 = [u64(0), u64(0), u64(0)]
    lib.mfcas_counters(ctx, *[ctypes.byref(c) for c in ec1])
    h2d, d2h = ec1[1].value - ec0[1].value, ec1[2].value - ec0[2].value         # bytes the library really moved each way
    assert h_out[(ne - 1) * stride:(ne - 1) * stride + blob_len].tobytes() == h_in[(ne - 1) * stride:(ne - 1) * stride + blob_len].tobytes()

    # ---- L-api: objects -> TaskDataStore.save_artifacts -> files on tmpfs -> load_artifacts -> objects,
    # pickle and filesystem included; a sample of the same arrays on EVERY rank (its own GPU, its own datastore root on
    # tmpfs), timed between barriers, aggregated like the other legs (all ranks' bytes / slowest rank's time)
    api = None
    api_t = [0.0, 0.0]
    # the object-level sample is sized per JOB (1024 artifacts = 17 GB in total at the default), not per rank: its host memory
    # (arrays, files on tmpfs, decoded copies) must not grow with the GPU count
    api_blobs = min(max(args.api_blobs // world, min(args.api_blobs, 64)), n)
    del d_packed, d_out                                   # the object-level leg brings its own context and arenas
    torch.cuda.empty_cache()
    if api_blobs > 0:
        import shutil
        import tempfile
        from metaflow_b200 import FlowDataStore, LocalStorage
        from metaflow_b200.codec import GpuCodec
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        root = tempfile.mkdtemp(prefix="mfcas_api_", dir=base)
        codec = GpuCodec(devices=[local_rank], wave_bytes=args.wave_mib << 20)
        try:
            arrays = [d_in[i * stride + len(prefix):i * stride + len(prefix) + nbytes].cpu().numpy().view(np.float32).copy() for i in range(api_blobs)]
            fds = FlowDataStore("BenchFlow", storage_impl=LocalStorage, ds_root=root, codec=codec)
            for attempt in range(2):                       # first pass warms the arenas
                w = fds.get_task_datastore("1", "work", str(attempt), attempt=0, mode="w")
                w.init_task()
                shutil.rmtree(os.path.join(root, "BenchFlow", "data"), ignore_errors=True)
                barrier()
                t0 = time.perf_counter()
                w.save_artifacts(((("a%d" % i), a) for i, a in enumerate(arrays)), len_hint=len(arrays))
                w.done()
                t1 = time.perf_counter()
                barrier()
                t1b = time.perf_counter()
                r = fds.get_task_datastore("1", "work", str(attempt))
                back = dict(r.load_artifacts(["a%d" % i for i in range(len(arrays))]))
                t2 = time.perf_counter()
            assert all(np.array_equal(back["a%d" % i], a) for i, a in enumerate(arrays)), "L-api round trip mismatch"
            api_t = [t1 - t0, t2 - t1b]
        finally:
            codec.close()
            shutil.rmtree(root, ignore_errors=True)

    # ---- reduce over ranks (max time), aggregate
    probe = torch.tensor([both_each], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(probe, op=dist.ReduceOp.MIN)      # the slowest rank bounds a barrier-to-barrier measurement
    both_each = float(probe[0])
    times = torch.tensor([t_save, t_load, e2e_save * 1e3, e2e_load * 1e3] + api_t + kms, dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_save, t_load, e2e_save_ms, e2e_load_ms = (float(x) for x in times[:4])
    api_t = [float(x) for x in times[4:6]]
    kms = [float(x) for x in times[6:]]
    if api_blobs > 0 and api_t[0] > 0 and api_t[1] > 0:
        nb = api_blobs * blob_len * world
        api = {"value": nb / (api_t[0] + api_t[1]) / 1e9, "unit": UNIT, "save_GBps": nb / api_t[0] / 1e9, "load_GBps": nb / api_t[1] / 1e9,
               "sample": "%d artifacts per rank through TaskDataStore.save_artifacts/load_artifacts, LocalStorage on tmpfs, pickle + files included" % api_blobs}
    total_raw = raw_bytes * world
    value = total_raw / ((t_save + t_load) * 1e-3) / 1e9
    e2e_value = total_raw / ((e2e_save_ms + e2e_load_ms) * 1e-3) / 1e9

    if rank == 0:
        peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                peak, peak_src = json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        names = ["sha1", "deflate", "compact", "inflate", "verify+foreign"]
        dominant = max(range(4), key=lambda i: kms[i])
        ratio = packed_bytes / raw_bytes
        # algorithmic bytes of the dominant kernel per launch (save N+C+20, load C+N)
        alg = {0: n * blob_len + 20 * n, 1: n * blob_len * (1 + ratio), 2: 2 * n * blob_len * ratio, 3: n * blob_len * (1 + ratio)}[dominant]
        # every kernel of the path against the same roofline (SURVEY 8d algorithmic bytes; SHA-1 only reads N)
        per_kernel = {}
        for k, nm in ((0, "sha1"), (1, "deflate"), (3, "inflate")):
            kb = {0: n * blob_len + 20 * n, 1: n * blob_len * (1 + ratio), 3: n * blob_len * (1 + ratio)}[k]
            kt = kms[k] / args.steps
            per_kernel[nm] = {"ms": kt, "GBps": kb / (kt * 1e-3) / 1e9 if kt > 0 else 0.0, "frac": (kb / (kt * 1e-3) / 1e9 / peak) if kt > 0 else 0.0}
        save_path_frac = (n * blob_len * (1 + ratio) + 20 * n) / ((t_save / args.steps) * 1e-3) / 1e9 / peak
        load_path_frac = (n * blob_len * (1 + ratio)) / ((t_load / args.steps) * 1e-3) / 1e9 / peak
        dom_ms = kms[dominant] / args.steps
        achieved = alg / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(REPO, "profiles", "dram_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            # measured with ncu on tj["blobs"] blobs of tj["blob_bytes"] bytes; every kernel streams each byte once, so the
            # traffic of a launch is linear in the raw bytes it processes (whatever the blob size)
            per_byte = tj.get("per_raw_byte", {}).get(names[dominant])
            traffic = per_byte * n * blob_len if per_byte else None
        cpu = None
        if not args.no_cpu_baseline:
            cores = usable_cores()
            per = max(1, args.ref_blobs_per_proc)
            cpu_reference_pass(cores, 1, 99)
            b, ts, tl = cpu_reference_pass(cores, per, 5_000_000)
            cpu = {"value": b / (ts + tl) / 1e9, "unit": UNIT, "cores": cores, "logical_cpus": os.cpu_count(), "kind": "port",
                   "sample": "%d blobs (%d procs x %d) through oracle/ref_path.py: hashlib.sha1 + gzip level 3 + gunzip + tmpfs CAS files" % (cores * per, cores, per),
                   "save_GBps": b / ts / 1e9, "load_GBps": b / tl / 1e9}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (t_save + t_load) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s (pickle blob %d B), save+load" % (args.workload, blob_len), "name": args.config,
                       "verified_blobs": verified,
                       "blobs_per_step_per_gpu": n, "blobs_total": n * args.steps * world, "sharding": "round-robin blob i -> rank i mod N, no collective",
                       "l2": "inputs %.1f GB per step >> 126 MB L2 (no flush needed)" % (n * blob_len / 1e9), "packed_ratio": ratio,
                       "e2e_sub_batch": ne, "wave_mib": args.wave_mib},
            "save_GBps": total_raw / (t_save * 1e-3) / 1e9, "load_GBps": total_raw / (t_load * 1e-3) / 1e9,
            "kernel_ms_per_step": {nm: k / args.steps for nm, k in zip(names, kms)},
            "roofline": {"bound": "hbm", "kernel": names[dominant], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "kernels": per_kernel, "save_path_frac": save_path_frac, "load_path_frac": load_path_frac},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // e2e_steps, "d2h_bytes_per_step": d2h // e2e_steps,
                    "save_GBps": total_raw / (e2e_save_ms * 1e-3) / 1e9, "load_GBps": total_raw / (e2e_load_ms * 1e-3) / 1e9,
                    "api": "mfcas_pack_batch_extents + mfcas_unpack_batch on page-locked host buffers",
                    # copies only, both directions busy on every rank at once: per-rank GB/s per direction, and the round-trip
                    # rate that alone would allow (save = N up | C down, load = C up | N down  ->  about both/2 per rank)
                    "copy_probe_GBps_each_way_per_rank": both_each, "copy_only_ceiling_GBps": world * both_each / 2.0,
                    "frac_of_copy_ceiling": e2e_value / (world * both_each / 2.0) if both_each > 0 else None},
            "api": api,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    lib.mfcas_free_pinned(h_in_p); lib.mfcas_free_pinned(h_pk_p); lib.mfcas_free_pinned(h_out_p)
    lib.mfcas_destroy(ctx)
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS), help="BASELINE.json config (the driver's default run is C2)")
    ap.add_argument("--blobs-per-step", type=int, default=0, help="0 = the config's own wave size")
    ap.add_argument("--verify-frac", type=float, default=0.01, help="fraction of every verified step's blobs checked against the oracle")
    ap.add_argument("--wave-mib", type=int, default=1024, help="bytes (in+out) per pipelined wave of the host-pointer API")
    ap.add_argument("--ref-blobs-per-proc", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--api-blobs", type=int, default=1024, help="artifacts of the whole job in the L-api (objects <-> tmpfs files) leg, split over the ranks (at least 64 each); 0 = skip")
    args = ap.parse_args()
    global ARRAY_ELEMS
    ARRAY_ELEMS = CONFIGS[args.config][0]
    if args.blobs_per_step <= 0:
        args.blobs_per_step = CONFIGS[args.config][1]
    args.workload = CONFIGS[args.config][2]
    args.ref_blobs_per_proc *= max(1, 4194304 // ARRAY_ELEMS)        # the CPU sample stays about 64 MiB per process
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
