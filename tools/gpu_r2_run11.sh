#!/bin/bash
# Round 2, run 11: decoder occupancy variants, e2e load with product-style staging (C3), copy-only ceiling probe.
mkdir -p gpurun_out
{
echo "=== pytest parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -2
for occ in 5 4; do
echo "=== bench C2 2048 dec_occ=$occ"; MFCAS_DEC_OCC=$occ timeout 1200 python bench.py --blobs-per-step 2048 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'], d['e2e'])"; tail -2 gpurun_out/c.err
done
echo "=== bench C3"; timeout 1200 python bench.py --config C3 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c.err | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['kernel_ms_per_step'], d['e2e'])"; tail -2 gpurun_out/c.err
} > gpurun_out/r2_run11.log 2>&1
tail -c 3500 gpurun_out/r2_run11.log
