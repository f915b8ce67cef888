#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
} > gpurun_out/tests_only.log 2>&1
tail -c 1500 gpurun_out/tests_only.log
