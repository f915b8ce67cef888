#!/bin/bash
# First GPU session: smoke, sanitizer on a tiny case, parity tests, short bench.
mkdir -p gpurun_out
{
echo "=== nvidia-smi"; nvidia-smi -L
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
echo "=== sanitizer (smoke)"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke 2>&1 | tail -40
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
echo "=== bench small"; timeout 900 python bench.py --steps 2 --warmup 1 --blobs-per-step 64 --ref-blobs-per-proc 1
} > gpurun_out/first_run.log 2>&1
tail -c 6000 gpurun_out/first_run.log
