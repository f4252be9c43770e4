#!/bin/bash
# 4-GPU: ticketed RS reduce (correctness + timing at TP=4 shapes), ring attention P2P path, bench N=4 fused + profile
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=4 timeout -k 10 240 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused4.log 2>&1; echo "fused4 rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch\|timeout" gpurun_out/fused4.log | cut -c1-700 | tail -12
NGPU=2 timeout -k 10 200 python tests/test_shardformer/test_ring_attention.py > gpurun_out/ring2.log 2>&1; echo "ring rc=$?"; tail -3 gpurun_out/ring2.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 420 $TR bench.py --gpus 4 --steps 2 --warmup 3 --no-e2e --profile gpurun_out/prof_n4_fused.txt > gpurun_out/b4_fused.log 2>&1; echo "bench4 fused rc=$?"; grep -a '"metric"' gpurun_out/b4_fused.log | cut -c1-600
head -12 gpurun_out/prof_n4_fused.txt | cut -c1-150
