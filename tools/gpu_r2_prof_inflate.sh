#!/bin/bash
# one `ncu --set full` capture of the chunk decoder alone (256 blobs), for the stall / pipe breakdown
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mf_inflate_chunk_kernel" -s 1 -c 1 -o gpurun_out/r02_inflate_prof -f python bench.py --blobs-per-step 256 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_inflate.log 2>&1
tail -2 gpurun_out/ncu_inflate.log | cut -c1-300; ls -la gpurun_out/r02_inflate_prof.ncu-rep
