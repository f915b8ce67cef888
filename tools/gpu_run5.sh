#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "=== phase timing (64 blobs)"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --blobs-per-step 64 --no-cpu-baseline 2>&1 | grep -E "phase" | cut -c1-900 | tail -3
echo "=== bench 512"; timeout 1200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline
} > gpurun_out/run5.log 2>&1
tail -c 6000 gpurun_out/run5.log
