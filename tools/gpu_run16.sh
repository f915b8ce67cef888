#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== bench default"; S=$(date +%s); timeout 1500 python bench.py; echo "elapsed $(( $(date +%s) - S )) s"
echo "=== phase profile"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] deflate" | tail -2
} > gpurun_out/run16.log 2>&1
tail -c 3000 gpurun_out/run16.log
