#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
} > gpurun_out/run19.log 2>&1
tail -c 3000 gpurun_out/run19.log
