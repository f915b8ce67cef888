#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
echo "=== default bench"; S=$(date +%s); timeout 1700 python bench.py 2> gpurun_out/bench_default.err; echo "elapsed $(( $(date +%s) - S )) s"; tail -2 gpurun_out/bench_default.err
} > gpurun_out/run25.log 2>&1
tail -c 5000 gpurun_out/run25.log
