#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu (decoder subset)"; timeout 1500 python -m pytest tests -m gpu -q -k "golden or edge or foreign or corruption or baseline" 2>&1 | tail -3
echo "=== bench"; timeout 1500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"
} > gpurun_out/run24.log 2>&1
tail -c 3000 gpurun_out/run24.log
