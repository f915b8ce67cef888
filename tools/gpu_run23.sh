#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== foreign decode"; timeout 600 python tools/foreign_decode.py 2>&1 | tail -3
} > gpurun_out/run23.log 2>&1
tail -c 3000 gpurun_out/run23.log
