#!/bin/bash
mkdir -p gpurun_out
{
echo "=== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -4
echo "=== bench 2048"; timeout 600 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
echo "=== sanitizer on a small bench"; timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python bench.py --blobs-per-step 16 --steps 1 --warmup 0 --no-cpu-baseline --api-blobs 0 2>&1 | grep -v "^$" | grep -v "^{" | head -60
} > gpurun_out/r2_san.log 2>&1
head -c 6000 gpurun_out/r2_san.log
