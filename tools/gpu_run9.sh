#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
for w in 512 1024 2048; do
echo "=== bench wave-mib $w"; timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib $w | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['e2e'])"
done
echo "=== timeline wave-mib 1024"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib 1024 2>&1 | grep -v "phase\]" | tail -130
} > gpurun_out/run9.log 2>&1
tail -c 3000 gpurun_out/run9.log
