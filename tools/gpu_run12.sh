#!/bin/bash
mkdir -p gpurun_out
{
echo "THP: $(cat /sys/kernel/mm/transparent_hugepage/enabled) defrag: $(cat /sys/kernel/mm/transparent_hugepage/defrag)"
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30
for spc in 32 16; do
echo "=== bench SHA_SPC=$spc"; MFCAS_SHA_SPC=$spc timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 --wave-mib 1024 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'], d['e2e']['save_GBps'], d['e2e']['load_GBps'])"
done
echo "=== api profile 64"; timeout 600 python tools/api_profile.py 64 2>&1 | tail -9
echo "=== api profile 256"; timeout 600 python tools/api_profile.py 256 2>&1 | tail -9
} > gpurun_out/run12.log 2>&1
tail -c 3000 gpurun_out/run12.log
