#!/bin/bash
# torchrun bench at 8 GPUs with the NUMA binding on (default); prints e2e / value / api
mkdir -p gpurun_out
{
nvidia-smi topo -m 2>/dev/null | grep -E "^GPU[0-9]" | awk '{print $1, $(NF-2), $(NF-1)}'
echo "=== torchrun N=8 bench (NUMA binding on)"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 64 2>gpurun_out/multi8.err | grep "^{" | tee gpurun_out/bench_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'],'api',d['api'])"; tail -2 gpurun_out/multi8.err
echo "=== torchrun N=8 bench (MFCAS_NUMA=0), e2e only"; MFCAS_NUMA=0 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/multi8b.err | grep "^{" | tee gpurun_out/bench_n8_nonuma.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'])"; tail -2 gpurun_out/multi8b.err
} > gpurun_out/r2_multi_8.log 2>&1
tail -c 4000 gpurun_out/r2_multi_8.log
