"""One-off probe of the GPU box: host cores/RAM, /dev/shm, PCIe pinned copy bandwidth, GPU info."""
import os, time, json, subprocess, shutil
import torch
info = {}
info["cpu_count"] = os.cpu_count()
info["sched_affinity"] = len(os.sched_getaffinity(0))
with open("/proc/meminfo") as f:
    mi = dict(l.split(":")[0:2] for l in f.read().splitlines() if ":" in l)
info["MemTotal"] = mi["MemTotal"].strip(); info["MemAvailable"] = mi["MemAvailable"].strip()
info["shm"] = shutil.disk_usage("/dev/shm")._asdict()
info["tmp"] = shutil.disk_usage("/tmp")._asdict()
info["lscpu"] = subprocess.run("lscpu | egrep 'Model name|Socket|Thread|Core|NUMA|Flags' | cut -c1-300", shell=True, capture_output=True, text=True).stdout
info["nvsmi"] = subprocess.run("nvidia-smi; nvidia-smi topo -m", shell=True, capture_output=True, text=True).stdout
info["ulimit_l"] = subprocess.run("ulimit -l", shell=True, capture_output=True, text=True).stdout
d = torch.device("cuda:0")
p = torch.cuda.get_device_properties(0)
info["gpu"] = dict(name=p.name, sms=p.multi_processor_count, mem=p.total_memory, cc=(p.major, p.minor))
for sz in (64 << 20, 1 << 30, 4 << 30):
    t0 = time.time(); h = torch.empty(sz, dtype=torch.uint8, pin_memory=True); info[f"pin_alloc_s_{sz>>20}MiB"] = time.time() - t0
    g = torch.empty(sz, dtype=torch.uint8, device=d)
    for name, fn in (("h2d", lambda: g.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(g, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); [fn() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        info[f"{name}_GBs_{sz>>20}MiB"] = 3 * sz / (e0.elapsed_time(e1) * 1e-3) / 1e9
    # bidirectional
    g2 = torch.empty(sz, dtype=torch.uint8, device=d); h2 = torch.empty(sz, dtype=torch.uint8, pin_memory=True)
    s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream(); torch.cuda.synchronize()
    t0 = time.time()
    with torch.cuda.stream(s1): g.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(g2, non_blocking=True)
    torch.cuda.synchronize(); info[f"bidir_GBs_each_{sz>>20}MiB"] = sz / (time.time() - t0) / 1e9
    del h, g, g2, h2
# pageable copy + host memcpy speed
import numpy as np
a = np.ones(1 << 30, dtype=np.uint8); b = np.empty_like(a)
t0 = time.time(); b[:] = a; info["host_memcpy_GBs_1thread"] = (1 << 30) / (time.time() - t0) / 1e9
import hashlib, zlib
blob = np.random.default_rng(0).standard_normal(4 << 20, dtype=np.float32).tobytes()
t0 = time.time(); hashlib.sha1(blob).digest(); info["sha1_GBs_1core"] = len(blob) / (time.time() - t0) / 1e9
t0 = time.time(); c = zlib.compress(blob, 3); info["zlib3_MBs_1core"] = len(blob) / (time.time() - t0) / 1e6; info["zlib3_ratio"] = len(c) / len(blob)
t0 = time.time(); zlib.decompress(c); info["inflate_MBs_1core"] = len(blob) / (time.time() - t0) / 1e6
t0 = time.time(); x = np.random.default_rng(1).standard_normal(16 << 20, dtype=np.float32); info["randn_MBs_1core"] = x.nbytes / (time.time() - t0) / 1e6
print(json.dumps(info, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(info, open("gpurun_out/probe_box.json", "w"), indent=1)
