#!/bin/bash
mkdir -p gpurun_out
{
echo "=== bench 2048 blobs/step"; S=$(date +%s); timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib 1024; echo "elapsed $(( $(date +%s) - S )) s"
} > gpurun_out/run14.log 2>&1
tail -c 3000 gpurun_out/run14.log
