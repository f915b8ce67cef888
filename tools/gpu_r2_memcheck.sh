#!/bin/bash
# compute-sanitizer memcheck over the GPU parity suite (final build): every kernel, every edge-size / corruption case
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_memcheck.log 2>&1
grep -v "^$" gpurun_out/r02_memcheck.log | tail -15
