#!/bin/bash
# parity first; the evidence run only if it is green
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3 > gpurun_out/r2_run16.log
cat gpurun_out/r2_run16.log
grep -q "failed\|error" gpurun_out/r2_run16.log && exit 1
{
echo "=== bench C2 2048"; timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"; tail -2 gpurun_out/c.err
} >> gpurun_out/r2_run16.log 2>&1
tail -3 gpurun_out/r2_run16.log
bash tools/gpu_r2_final_profile.sh
