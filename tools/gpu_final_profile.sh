#!/bin/bash
# Round-end evidence: GPU tests, default bench, reference arm, ncu launch list of the bench command, ncu --set full of the three hot kernels.
mkdir -p gpurun_out
{
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke
echo "=== default bench"; S=$(date +%s); timeout 1700 python bench.py 2> gpurun_out/bench_default.err; echo "elapsed $(( $(date +%s) - S )) s"; tail -2 gpurun_out/bench_default.err
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null
echo "=== api profile 256"; timeout 600 python tools/api_profile.py 256 2>&1 | tail -6
echo "=== ncu launch list (same command, one step)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mf_ -c 300 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log | cut -c1-300
echo "=== ncu full: sha1 / deflate / inflate (256 blobs per launch: traffic is linear in blobs)"
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:"mf_sha1_kernel|mf_deflate_kernel|mf_inflate_chunk_kernel" -s 3 -c 3 -o gpurun_out/r01_final_prof -f python bench.py --blobs-per-step 256 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/r01_final_prof.ncu-rep
} > gpurun_out/final.log 2>&1
tail -c 7000 gpurun_out/final.log
