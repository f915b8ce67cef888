#!/bin/bash
# Round 2, run 3: encoder v2 after the sampling fix: parity, phase profile, C2 bench, ncu full capture of the encoder.
mkdir -p gpurun_out
{
echo "=== pytest parity"; S=$(date +%s); timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -30; echo "elapsed $(( $(date +%s) - S )) s"
echo "=== phase profile v2"; MFCAS_PHASE_TIMING=1 timeout 600 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] deflate" | tail -2
echo "=== bench C2 v2"; timeout 1200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c2v2.err | tee gpurun_out/bench_c2_v2.json | cut -c1-2600; tail -5 gpurun_out/c2v2.err
echo "=== ncu full deflate2 (64 blobs)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:mf_deflate2_kernel -s 1 -c 1 -o gpurun_out/r02_deflate2 -f python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_d2.log 2>&1; tail -2 gpurun_out/ncu_d2.log | cut -c1-300
ls -la gpurun_out/r02_deflate2.ncu-rep
echo "=== foreign decode"; timeout 300 python tools/foreign_decode.py
echo "=== pytest shapes"; S=$(date +%s); timeout 1500 python -m pytest tests/test_gpu_shapes.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -30; echo "elapsed $(( $(date +%s) - S )) s"
} > gpurun_out/r2_run3.log 2>&1
tail -c 9000 gpurun_out/r2_run3.log
