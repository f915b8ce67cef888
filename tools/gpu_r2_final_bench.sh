#!/bin/bash
mkdir -p gpurun_out
{
echo "=== bench.py (default)"; s=$(date +%s); timeout 1200 python bench.py 2>gpurun_out/bench_default.err | tail -1 | tee gpurun_out/bench_default.json | cut -c1-6000; echo "wall $(( $(date +%s) - s )) s"
} > gpurun_out/r02_final_bench.log 2>&1
tail -c 7000 gpurun_out/r02_final_bench.log
