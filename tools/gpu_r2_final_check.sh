#!/bin/bash
# what the driver runs at round end, on the final build: smoke, the GPU suite, both bench arms with default flags
mkdir -p gpurun_out
{
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== pytest -m gpu"; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -v "^$" | tail -4
echo "=== bench.py --impl reference"; timeout 900 python bench.py --impl reference 2>/dev/null | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-900
echo "=== bench.py (default)"; /usr/bin/time -v timeout 1200 python bench.py 2>gpurun_out/bench_default.err | tail -1 | tee gpurun_out/bench_default.json | cut -c1-6000; grep -E "Elapsed|Maximum resident" gpurun_out/bench_default.err
} > gpurun_out/r02_final_check.log 2>&1
tail -c 8000 gpurun_out/r02_final_check.log
