#!/bin/bash
# 8-GPU: validate the fused kernels at world=8 + bench N=8
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=8 timeout -k 10 200 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused8c.log 2>&1; echo "fused8 rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch\|timeout" gpurun_out/fused8c.log | cut -c1-700 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 300 $TR bench.py --gpus 8 --steps 2 --warmup 3 --profile gpurun_out/prof_n8_c.txt > gpurun_out/b8c_fused.log 2>&1; echo "bench8 rc=$?"; grep -a '"metric"' gpurun_out/b8c_fused.log | cut -c1-1800
head -8 gpurun_out/prof_n8_c.txt | cut -c1-150
