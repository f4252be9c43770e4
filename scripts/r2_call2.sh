#!/bin/bash
# round-2 call 2 (2 GPUs): RS variants + fused GEMM+AR at TP=2, bench harness (both arms, both rows) on a cut-down model
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
NGPU=2 timeout 600 python tests/test_parallel/test_fused_comm.py > gpurun_out/c2_fused2.log 2>&1
echo "fused2 rc=$?" >> gpurun_out/c2_fused2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 3 --warmup 3 --layers 8 > gpurun_out/c2_bench_ours.log 2>&1
echo "bench_ours rc=$?" >> gpurun_out/c2_bench_ours.log
CB200_REF_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --impl reference --gpus 2 --steps 3 --warmup 3 --layers 8 > gpurun_out/c2_bench_ref.log 2>&1
echo "bench_ref rc=$?" >> gpurun_out/c2_bench_ref.log
grep -E "FUSED_TIMING|RS_TUNING|rc=|Error" gpurun_out/c2_fused2.log | tail -12
tail -c 1500 gpurun_out/c2_bench_ours.log; echo; tail -c 2500 gpurun_out/c2_bench_ref.log
