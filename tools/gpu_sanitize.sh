#!/bin/bash
mkdir -p gpurun_out
{
echo "=== compute-sanitizer memcheck (subset)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "edge or carried or corruption or golden or distributions or foreign or small_blobs or contract" 2>&1 | tail -12
echo "=== compute-sanitizer racecheck (foreign streams: shared-memory tables, warp CRC)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "foreign" 2>&1 | tail -12
} > gpurun_out/sanitize.log 2>&1
tail -c 4000 gpurun_out/sanitize.log
