#!/bin/bash
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60
echo "=== bench 64"; timeout 900 python bench.py --steps 2 --warmup 1 --blobs-per-step 64 --ref-blobs-per-proc 1
echo "=== bench 512 (default-size wave)"; timeout 1200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 1 --blobs-per-step 32 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
echo "=== ncu full: deflate + inflate"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mf_deflate_kernel|mf_inflate_chunk_kernel" -s 2 -c 2 -o gpurun_out/r01_prof -f python bench.py --steps 1 --warmup 1 --blobs-per-step 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
} > gpurun_out/run2.log 2>&1
tail -c 8000 gpurun_out/run2.log
