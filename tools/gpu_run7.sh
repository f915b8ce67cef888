#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_sha1 -s 1 -c 1 -o gpurun_out/r01_sha -f python tools/sha_probe.py 2>&1 | tail -3
} > gpurun_out/run7.log 2>&1
tail -c 2000 gpurun_out/run7.log
