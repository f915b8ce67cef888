#!/bin/bash
# Round-2 evidence for profiles/: ncu launch list of the default bench command (one step), `ncu --set full` of the three
# hot kernels (256 blobs per launch: the kernels stream every blob once, so DRAM traffic is linear in blobs), foreign
# decoder launch list, smoke.
mkdir -p gpurun_out
{
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
echo "=== ncu launch list (default command, one step)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mf_ -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-300
echo "=== ncu full: sha1 / deflate2 / inflate (256 blobs per launch)"
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:"mf_sha1_kernel|mf_deflate2_kernel|mf_inflate_chunk_kernel" -s 3 -c 3 -o gpurun_out/r02_final_prof -f python bench.py --blobs-per-step 256 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/r02_final_prof.ncu-rep
echo "=== foreign decode launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mf_ -c 200 --csv --log-file gpurun_out/r02_foreign_launches.csv python tools/foreign_decode.py > gpurun_out/ncu_foreign.log 2>&1; tail -3 gpurun_out/ncu_foreign.log
echo "=== foreign decode (no profiler)"; timeout 300 python tools/foreign_decode.py
echo "=== api profile 512"; timeout 600 python tools/api_profile.py 512 2>&1 | tail -6
echo "=== C3 / C4 bench lines"
timeout 900 python bench.py --config C3 --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/bench_c3.json | cut -c1-2000
timeout 900 python bench.py --config C4 --steps 2 --warmup 1 2>/dev/null | tee gpurun_out/bench_c4.json | cut -c1-2000
} > gpurun_out/r02_final.log 2>&1
tail -c 7000 gpurun_out/r02_final.log
