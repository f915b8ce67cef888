#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== phase (64)"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --blobs-per-step 64 --no-cpu-baseline 2>&1 | grep -E "phase\] deflate" | cut -c1-600 | tail -2
echo "=== bench (default 1024 x 2 steps)"; timeout 1500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline
} > gpurun_out/run9.log 2>&1
tail -c 5000 gpurun_out/run9.log
