#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== C3 small blobs"; timeout 900 python tools/c3_small_blobs.py 65536 2>&1 | tail -6
echo "=== api profile 64"; timeout 600 python tools/api_profile.py 64 2>&1 | tail -3
} > gpurun_out/run18.log 2>&1
tail -c 3000 gpurun_out/run18.log
