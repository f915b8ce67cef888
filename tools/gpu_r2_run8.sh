#!/bin/bash
# Round 2, run 8: default bench (3072 blobs/step, API leg), reference arm, parity tests, launch list of the default command.
mkdir -p gpurun_out
{
echo "=== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
echo "=== phase profile v2"; MFCAS_PHASE_TIMING=1 timeout 600 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] deflate" | tail -2 | head -1
echo "=== default bench"; S=$(date +%s); timeout 1700 python bench.py 2> gpurun_out/bench_default.err | tee gpurun_out/bench_default.json | cut -c1-3500; echo "elapsed $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench_default.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-900
} > gpurun_out/r2_run8.log 2>&1
tail -c 6000 gpurun_out/r2_run8.log
