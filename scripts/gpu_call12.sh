#!/bin/bash
# 2-GPU: CTA-pair fused kernels (correctness + timing), N=2 bench + profile, N=1 bench
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=2 timeout -k 10 420 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused2b.log 2>&1; echo "fused rc=$?"; grep -a "FUSED_\|Error\|error" gpurun_out/fused2b.log | cut -c1-600 | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --profile gpurun_out/prof_n2_fused_2cta.txt > gpurun_out/b2c_fused.log 2>&1; echo "bench2 rc=$?"; grep -a '"metric"' gpurun_out/b2c_fused.log | cut -c1-1400
head -14 gpurun_out/prof_n2_fused_2cta.txt | cut -c1-150
CUDA_VISIBLE_DEVICES=0 timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-e2e --profile gpurun_out/prof_n1_b.txt > gpurun_out/b1_b.log 2>&1; echo "bench1 rc=$?"; grep -a '"metric"' gpurun_out/b1_b.log | cut -c1-700
head -16 gpurun_out/prof_n1_b.txt | cut -c1-150
