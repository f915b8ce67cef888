#!/bin/bash
mkdir -p gpurun_out
{
for w in 512; do
echo "=== timeline wave-mib $w"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib $w 2>&1 | grep -v "phase\]" | tail -230
done
} > gpurun_out/run10.log 2>&1
tail -c 3000 gpurun_out/run10.log
