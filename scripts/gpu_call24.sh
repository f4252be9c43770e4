#!/bin/bash
# 2-GPU: functional check of the mixed dispatch (fused AG->GEMM + cuBLAS/NCCL GEMM->RS for comm-bound shapes)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
CB200_FUSED_RS_MIN_K=100000 timeout -k 10 200 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > gpurun_out/b2_mixed.log 2>&1; echo "bench2 rc=$?"; grep -a '"metric"' gpurun_out/b2_mixed.log | cut -c1-1500; grep -a "Error\|error" gpurun_out/b2_mixed.log | head -5
