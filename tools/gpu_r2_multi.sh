#!/bin/bash
# usage: gpu_r2_multi.sh N   -- torchrun bench at N GPUs (NUMA on / off), topology, two-device context test
N=${1:-2}
mkdir -p gpurun_out
{
nvidia-smi -L | head -8; nvidia-smi topo -m 2>/dev/null | head -14
echo "=== 2-device context test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_device" 2>&1 | tail -2
echo "=== torchrun N=$N bench (NUMA binding on)"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 128 2>gpurun_out/multi.err | tee gpurun_out/bench_n${N}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'],'api',d['api'])"; tail -2 gpurun_out/multi.err
echo "=== torchrun N=$N bench (MFCAS_NUMA=0)"; MFCAS_NUMA=0 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --api-blobs 0 2>gpurun_out/multi.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e'])"; tail -2 gpurun_out/multi.err
} > gpurun_out/r2_multi_$N.log 2>&1
tail -c 5000 gpurun_out/r2_multi_$N.log
