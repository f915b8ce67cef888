#!/bin/bash
mkdir -p gpurun_out
{
echo "=== compute-sanitizer memcheck (subset)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "edge or carried or corruption or golden_raw or distributions or foreign or small_blobs" 2>&1 | tail -25
echo "rc=$?"
echo "=== compute-sanitizer racecheck (edge sizes + carried)"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "edge or carried" 2>&1 | tail -25
} > gpurun_out/sanitize.log 2>&1
tail -c 4000 gpurun_out/sanitize.log
