#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "=== bench"; timeout 1500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['config']['packed_ratio'], d['kernel_ms_per_step'], d['e2e']['save_GBps'], d['e2e']['load_GBps'])"
echo "=== phase profile"; MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] sha1" | tail -2
echo "=== foreign decode"; timeout 600 python tools/foreign_decode.py 2>&1 | tail -3
echo "=== SHA rol30 on the FMA pipe: tests"; MFCAS_SHA_FMA=1 timeout 900 python -m pytest tests -m gpu -q -k "golden or edge or small_blobs or baseline or striped" 2>&1 | tail -2
echo "=== SHA rol30 on the FMA pipe: bench"; MFCAS_SHA_FMA=1 timeout 1500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['save_GBps'], d['load_GBps'], d['kernel_ms_per_step'])"
echo "=== SHA rol30 on the FMA pipe: phase"; MFCAS_SHA_FMA=1 MFCAS_PHASE_TIMING=1 timeout 900 python bench.py --blobs-per-step 64 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>&1 | grep "phase\] sha1" | tail -2
} > gpurun_out/run22.log 2>&1
tail -c 3000 gpurun_out/run22.log
