#!/bin/bash
# 8-GPU: fused kernels at TP=8 (correctness + timing), MoE fused EP at 8 ranks, bench fused vs nccl backend
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=8 timeout -k 10 300 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused8.log 2>&1; echo "fused8 rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch" gpurun_out/fused8.log | cut -c1-700 | tail -14
NGPU=8 timeout -k 10 200 python tests/test_moe/test_moe_ops.py > gpurun_out/moe_ep8.log 2>&1; echo "moe8 rc=$?"; tail -3 gpurun_out/moe_ep8.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 600 $TR bench.py --gpus 8 --steps 2 --warmup 3 --no-e2e --profile gpurun_out/prof_n8_fused.txt > gpurun_out/b8_fused.log 2>&1; echo "bench8 fused rc=$?"; grep -a '"metric"' gpurun_out/b8_fused.log | cut -c1-700
head -16 gpurun_out/prof_n8_fused.txt | cut -c1-150
timeout -k 10 600 $TR bench.py --gpus 8 --steps 2 --warmup 3 --no-e2e --comm-backend nccl --profile gpurun_out/prof_n8_nccl.txt > gpurun_out/b8_nccl.log 2>&1; echo "bench8 nccl rc=$?"; grep -a '"metric"' gpurun_out/b8_nccl.log | cut -c1-700
head -12 gpurun_out/prof_n8_nccl.txt | cut -c1-150
