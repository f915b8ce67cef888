#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
for v in 0 1; do
echo "=== bench C2 2048 dec_crc=$v"; MFCAS_DEC_CRC=$v timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
done
} > gpurun_out/r2_run15.log 2>&1
tail -c 3000 gpurun_out/r2_run15.log
