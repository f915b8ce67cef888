#!/bin/bash
# GPU test suite, a short bench run, the SHA-1 unit profile and the foreign-stream decode timing.
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== bench"; timeout 1500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['config']['packed_ratio'], d['kernel_ms_per_step'], d['e2e']['save_GBps'], d['e2e']['load_GBps'])"
echo "=== phase profile"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] " | tail -4
echo "=== foreign decode"; timeout 600 python tools/foreign_decode.py 2>&1 | tail -3
} > gpurun_out/run22.log 2>&1
tail -c 3000 gpurun_out/run22.log
