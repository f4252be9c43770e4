#!/bin/bash
# 2-GPU: bench ours (fused / nccl) + reference arm
export PYTHONPATH=$PWD
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 300 $TR bench.py --gpus 2 --steps 2 --warmup 3 --layers 4 --no-e2e > gpurun_out/b2_quick.log 2>&1; echo "quick rc=$?"; tail -2 gpurun_out/b2_quick.log | cut -c1-600
timeout -k 10 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/b2_fused.log 2>&1; echo "fused rc=$?"; tail -1 gpurun_out/b2_fused.log | cut -c1-1200
timeout -k 10 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --comm-backend nccl --no-e2e > gpurun_out/b2_nccl.log 2>&1; echo "nccl rc=$?"; tail -1 gpurun_out/b2_nccl.log | cut -c1-1200
timeout -k 10 900 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/b2_ref.log 2>&1; echo "ref rc=$?"; grep -a "reference arm\|impl" gpurun_out/b2_ref.log | cut -c1-1500 | tail -8
