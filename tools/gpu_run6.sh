#!/bin/bash
mkdir -p gpurun_out
{
echo "=== phase+timeline (512 blobs, 1 step)"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -E "phase|timeline|^ +[0-9]+:" | cut -c1-400 | tail -60
} > gpurun_out/run6.log 2>&1
tail -c 9000 gpurun_out/run6.log
