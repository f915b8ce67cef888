#!/bin/bash
# Round 2, run 2: first run of the member-owned encoder (mf_deflate2): parity tests, phase profile, C2 bench, A/B against v1.
mkdir -p gpurun_out
{
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== pytest parity (new encoder)"; S=$(date +%s); timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -30; echo "elapsed $(( $(date +%s) - S )) s"
echo "=== phase profile v2"; MFCAS_PHASE_TIMING=1 timeout 600 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] deflate" | tail -3
echo "=== bench C2 v2"; timeout 1200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c2v2.err | tee gpurun_out/bench_c2_v2.json | cut -c1-2500; tail -5 gpurun_out/c2v2.err
echo "=== pytest shapes"; S=$(date +%s); timeout 1500 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -30; echo "elapsed $(( $(date +%s) - S )) s"
} > gpurun_out/r2_run2.log 2>&1
tail -c 8000 gpurun_out/r2_run2.log
