#!/bin/bash
# 4-GPU: final multi-GPU validation of the fused kernels (world=4) + bench N=4 + e2e
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=4 timeout -k 10 240 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused4c.log 2>&1; echo "fused4 rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch\|timeout" gpurun_out/fused4c.log | cut -c1-700 | tail -10
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 420 $TR bench.py --gpus 4 --steps 2 --warmup 3 --profile gpurun_out/prof_n4_c.txt > gpurun_out/b4c_fused.log 2>&1; echo "bench4 rc=$?"; grep -a '"metric"' gpurun_out/b4c_fused.log | cut -c1-1800
head -6 gpurun_out/prof_n4_c.txt | cut -c1-150
