#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mf_ -c 100 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"mf_inflate_chunk_kernel" -s 1 -c 1 -o gpurun_out/r01_final_inflate -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log | cut -c1-200
} > gpurun_out/run12.log 2>&1
tail -c 1500 gpurun_out/run12.log
