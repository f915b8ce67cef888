#!/bin/bash
mkdir -p gpurun_out
{
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mf_deflate_kernel|mf_inflate_chunk_kernel" -s 2 -c 2 -o gpurun_out/r01b_prof -f python bench.py --steps 1 --warmup 1 --blobs-per-step 32 --no-cpu-baseline 2>&1 | tail -3
} > gpurun_out/run10.log 2>&1
tail -c 1500 gpurun_out/run10.log
