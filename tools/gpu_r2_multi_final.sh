#!/bin/bash
# final build on N GPUs: two-device context test + the torchrun bench the driver launches
N=${1:-2}
mkdir -p gpurun_out
{
echo "=== 2-device context test"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_device" 2>&1 | tail -2
echo "=== torchrun N=$N bench"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --api-blobs 1024 2>gpurun_out/multi.err | grep "^{" | tee gpurun_out/bench_final_n${N}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'save',d['save_GBps'],'load',d['load_GBps'],'kernels',d['kernel_ms_per_step'],'e2e',d['e2e'],'api',d['api'])"; tail -2 gpurun_out/multi.err
} > gpurun_out/r2_multi_final_$N.log 2>&1
tail -c 4000 gpurun_out/r2_multi_final_$N.log
