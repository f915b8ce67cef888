#!/bin/bash
mkdir -p gpurun_out
{
echo "=== phase timing (64 blobs)"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --blobs-per-step 64 --no-cpu-baseline 2>&1 | grep -E "phase" | cut -c1-900 | tail -4
} > gpurun_out/run4.log 2>&1
tail -c 6000 gpurun_out/run4.log
