#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "ratios|passed|failed|Error|assert" | head -30
echo "=== bench 1024"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib 1024 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['packed_ratio'], d['kernel_ms_per_step'], d['e2e']['save_GBps'], d['e2e']['load_GBps'])"
echo "=== bench 2048"; timeout 900 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib 1024 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['packed_ratio'], d['kernel_ms_per_step'], d['e2e']['save_GBps'], d['e2e']['load_GBps'])"
echo "=== phase profile"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] deflate" | tail -2
} > gpurun_out/run15.log 2>&1
tail -c 3000 gpurun_out/run15.log
