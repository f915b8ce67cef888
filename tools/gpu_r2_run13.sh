#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
for g in 8 16 32 64 4; do
echo "=== bench C2 2048 dec_group=$g"; MFCAS_DEC_GROUP=$g timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
done
echo "=== sha variant 4: parity"; MFCAS_SHA_VARIANT=4 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
echo "=== bench C2 2048 sha variant 4"; MFCAS_SHA_VARIANT=4 timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
echo "=== bench C2 default (3072)"; timeout 1200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'], d['e2e'])"; tail -2 gpurun_out/c.err
} > gpurun_out/r2_run13.log 2>&1
tail -c 3000 gpurun_out/r2_run13.log
