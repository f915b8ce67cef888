#!/bin/bash
# Round 2, run 6: encoder with exact statistics on rebuild; full GPU suite; C2 at 2048 and 3072 blobs per step.
mkdir -p gpurun_out
{
echo "=== pytest gpu (all)"; S=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -8; echo "elapsed $(( $(date +%s) - S )) s"
echo "=== phase profile v2"; MFCAS_PHASE_TIMING=1 timeout 600 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] deflate" | tail -2 | head -1
for b in 2048 3072; do
echo "=== bench C2 $b blobs/step"; timeout 1200 python bench.py --blobs-per-step $b --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c2.err | tee gpurun_out/bench_c2_$b.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['config']['packed_ratio'], d['kernel_ms_per_step'], d['roofline']['kernels'], d['e2e']['value'])"; tail -3 gpurun_out/c2.err
done
echo "=== foreign decode"; timeout 300 python tools/foreign_decode.py
} > gpurun_out/r2_run6.log 2>&1
tail -c 5000 gpurun_out/r2_run6.log
