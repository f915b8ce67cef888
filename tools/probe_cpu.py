import os, time, zlib, multiprocessing as mp, numpy as np, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print("cpu_count", os.cpu_count(), "usable", bench.usable_cores())
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu.stat"):
    try: print(p, open(p).read().strip()[:200].replace("\n", " | "))
    except OSError as e: print(p, "ERR", e)
blob = np.random.default_rng(0).standard_normal(1 << 20, dtype=np.float32).tobytes()
def work(_):
    t = time.perf_counter(); zlib.compress(blob, 3); return time.perf_counter() - t
for n in (1, 8, 16, 32, 64, 128):
    with mp.get_context("fork").Pool(n) as pool:
        t0 = time.perf_counter(); ts = pool.map(work, range(n * 2)); el = time.perf_counter() - t0
    print("procs", n, "aggregate MB/s", round(2 * n * len(blob) / el / 1e6, 1), "mean per-call s", round(sum(ts) / len(ts), 3))
