#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
echo "=== bench C2 2048"; timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
echo "=== bench C2 2048 group 8"; MFCAS_DEC_GROUP=8 timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
echo "=== bench C2 2048 dec_crc 0"; MFCAS_DEC_CRC=0 timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
echo "=== pytest shapes"; timeout 1500 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
echo "=== bench C3"; timeout 600 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'], d['e2e']['value'])"; tail -2 gpurun_out/c.err
echo "=== bench C4"; timeout 600 python bench.py --config C4 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'], d['e2e']['value'])"; tail -2 gpurun_out/c.err
} > gpurun_out/r2_run14.log 2>&1
tail -c 3000 gpurun_out/r2_run14.log
