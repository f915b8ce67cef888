#!/bin/bash
mkdir -p gpurun_out
{
for f in 0 2 4 6; do
echo "=== MFCAS_DBG=$f"; MFCAS_DBG=$f timeout 600 python - <<'PY'
import ctypes, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from metaflow_b200 import capi
lib = capi.load()
ctx = ctypes.c_void_p(lib.mfcas_create((ctypes.c_int*1)(0), 1, 1<<30, 0))
n = 1024; L = 16777379; stride = (L + 255)//256*256 + 256
cap = lib.mfcas_pack_bound(L); cs = (cap+255)//256*256+256
d_in = torch.randint(0, 256, (n*stride,), dtype=torch.uint8, device="cuda")
d_in.view(torch.float32)[:] = torch.randn(n*stride//4, device="cuda")
d_pk = torch.empty(n*cs, dtype=torch.uint8, device="cuda")
u64, vp = ctypes.c_uint64, ctypes.c_void_p
src = (vp*n)(*[d_in.data_ptr()+i*stride for i in range(n)]); slen = (u64*n)(*([L]*n))
pk = (vp*n)(*[d_pk.data_ptr()+i*cs for i in range(n)]); pcap = (u64*n)(*([cap]*n)); plen = (u64*n)()
st = (ctypes.c_int32*n)(); sha = ctypes.create_string_buffer(20*n)
for it in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    rc = lib.mfcas_pack_device(ctx, 0, n, src, slen, pk, pcap, plen, sha, st)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    ms = (ctypes.c_float*8)(); lib.mfcas_last_kernel_ms(ctx, ms)
print("rc", rc, "pack %.1f ms  sha %.1f deflate %.1f gather %.3f" % (dt*1e3, ms[0], ms[1], ms[2]))
PY
done
} > gpurun_out/run21.log 2>&1
tail -c 2000 gpurun_out/run21.log
