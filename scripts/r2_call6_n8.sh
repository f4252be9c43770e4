#!/bin/bash
# round-2 call 6 (8 GPUs): TP=8 fused comm kernels (streamed in-switch GEMM->RS, fused GEMM+AR) vs NCCL+cuBLAS, then the
# N=8 bench (tp8+sp row and zero1(dp8) row) with a kernel breakdown of the tp row
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
NGPU=8 timeout 600 python tests/test_parallel/test_fused_comm.py > gpurun_out/c6_fused8.log 2>&1
echo "fused8 rc=$?" >> gpurun_out/c6_fused8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
  bench.py --gpus 8 --steps 5 --warmup 3 --profile gpurun_out/c6_prof_n8_tp.txt > gpurun_out/c6_bench_n8.log 2>&1
echo "bench8 rc=$?" >> gpurun_out/c6_bench_n8.log
grep -E "FUSED_TIMING|RS_TUNING|FUSED_STATS|rc=|Error|error|timeout" gpurun_out/c6_fused8.log | cut -c1-700 | tail -24
tail -c 3500 gpurun_out/c6_bench_n8.log
