#!/bin/bash
# Round 2, run 1: the new shape tests, SHA-1 consumer variants, C2/C3/C4 bench lines.
mkdir -p gpurun_out
{
echo "=== pytest gpu"; S=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -25; echo "elapsed $(( $(date +%s) - S )) s"
for v in 0 1 3; do
echo "=== sha variant $v"; MFCAS_SHA_VARIANT=$v MFCAS_PHASE_TIMING=1 timeout 600 python bench.py --blobs-per-step 256 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep -E "phase\] sha1|kernel_ms" | tail -2 | cut -c1-400
done
echo "=== bench C2"; timeout 1500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/c2.err | tee gpurun_out/bench_c2.json | cut -c1-1500
echo "=== bench C3"; timeout 900 python bench.py --config C3 --steps 2 --warmup 1 --api-blobs 0 2>gpurun_out/c3.err | tee gpurun_out/bench_c3.json | cut -c1-1500; tail -3 gpurun_out/c3.err
echo "=== bench C4"; timeout 900 python bench.py --config C4 --steps 2 --warmup 1 --api-blobs 0 2>gpurun_out/c4.err | tee gpurun_out/bench_c4.json | cut -c1-1500; tail -3 gpurun_out/c4.err
echo "=== foreign decode"; timeout 300 python tools/foreign_decode.py
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; cat /sys/fs/cgroup/cpu.max; df -h /dev/shm | tail -1; numactl -H 2>/dev/null | head -4; lscpu | grep -i numa
} > gpurun_out/r2_run1.log 2>&1
tail -c 6000 gpurun_out/r2_run1.log
