#!/bin/bash
# 2-GPU: fused kernels with the BK=128 main loop + staggered RS (correctness + timing), N=2 bench, N=1 bench native vs cublas
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=2 timeout -k 10 300 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused2e.log 2>&1; echo "fused2 rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch\|timeout" gpurun_out/fused2e.log | cut -c1-700 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --profile gpurun_out/prof_n2_e.txt > gpurun_out/b2f_fused.log 2>&1; echo "bench2 rc=$?"; grep -a '"metric"' gpurun_out/b2f_fused.log | cut -c1-400
head -8 gpurun_out/prof_n2_e.txt | cut -c1-150
CUDA_VISIBLE_DEVICES=0 CB200_GEMM_BACKEND=native timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-e2e --profile gpurun_out/prof_n1_native.txt > gpurun_out/b1_native.log 2>&1; echo "bench1 native rc=$?"; grep -a '"metric"' gpurun_out/b1_native.log | cut -c1-330
head -7 gpurun_out/prof_n1_native.txt | cut -c1-150
CUDA_VISIBLE_DEVICES=1 CB200_GEMM_BACKEND=cublas timeout -k 10 600 python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/b1_cublas.log 2>&1; echo "bench1 cublas rc=$?"; grep -a '"metric"' gpurun_out/b1_cublas.log | cut -c1-330
