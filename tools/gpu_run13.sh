#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
for i in 1 2; do
echo "=== bench"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib 1024 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'], d['e2e']['save_GBps'], d['e2e']['load_GBps'])"
done
} > gpurun_out/run13.log 2>&1
tail -c 3000 gpurun_out/run13.log
