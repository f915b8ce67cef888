#!/bin/bash
mkdir -p gpurun_out
{
echo "=== full default bench (with cpu baseline)"; /usr/bin/time -v timeout 1700 python bench.py 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err | head -2; grep -E "Elapsed|Maximum resident" gpurun_out/bench_default.err
echo "=== reference arm"; /usr/bin/time -v timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref.err; grep -E "Elapsed" gpurun_out/bench_ref.err
} > gpurun_out/run11.log 2>&1
tail -c 6000 gpurun_out/run11.log
