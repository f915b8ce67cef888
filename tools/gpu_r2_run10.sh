#!/bin/bash
# Round 2, run 10: tiny chunks stored + full-SM encoder grid (C3/C4), lean decoder reader.
mkdir -p gpurun_out
{
echo "=== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -3
for cfg in C2 C3 C4; do
echo "=== bench $cfg"; timeout 1200 python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['config']['packed_ratio'], d['kernel_ms_per_step'], d['e2e']['value'], d['e2e']['save_GBps'], d['e2e']['load_GBps'])"; tail -2 gpurun_out/c.err
done
} > gpurun_out/r2_run10.log 2>&1
tail -c 3000 gpurun_out/r2_run10.log
