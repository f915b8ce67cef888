#!/bin/bash
mkdir -p gpurun_out
{
nvidia-smi -L
echo "=== 2-device context test"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_device or golden_raw" 2>&1 | tail -5
echo "=== torchrun N=2 bench"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -3
echo "=== torchrun N=2 reference arm"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -2
} > gpurun_out/run_2gpu.log 2>&1
tail -c 6000 gpurun_out/run_2gpu.log
