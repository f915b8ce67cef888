#!/bin/bash
mkdir -p gpurun_out
{
echo "=== full default bench (with cpu baseline)"; S=$(date +%s); timeout 1700 python bench.py 2> gpurun_out/bench_default.err; echo "elapsed $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench_default.err
echo "=== reference arm"; S=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref.err; echo "elapsed $(( $(date +%s) - S )) s"; tail -3 gpurun_out/bench_ref.err
} > gpurun_out/run11.log 2>&1
tail -c 6000 gpurun_out/run11.log
