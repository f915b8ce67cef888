#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
echo "=== api profile 64"; timeout 600 python tools/api_profile.py 64 2>&1 | tail -12
echo "=== api profile 256"; timeout 600 python tools/api_profile.py 256 2>&1 | tail -12
} > gpurun_out/run11.log 2>&1
tail -c 4000 gpurun_out/run11.log
