#!/bin/bash
# round-2 call 1: first hardware run of the tcgen05 flash fwd + the streamed in-switch GEMM->RS (2 GPUs)
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_gpu.txt 2>&1
CB200_FLASH_NATIVE=1 timeout 180 python scripts/bench_flash_attn.py > gpurun_out/c1_flash_bench.log 2>&1
echo "flash_bench rc=$?" >> gpurun_out/c1_flash_bench.log
CB200_RS_STREAM=1 NGPU=2 timeout 420 python tests/test_parallel/test_fused_comm.py > gpurun_out/c1_fused2_stream.log 2>&1
echo "fused2_stream rc=$?" >> gpurun_out/c1_fused2_stream.log
tail -5 gpurun_out/c1_flash_test.log; tail -5 gpurun_out/c1_flash_bench.log; tail -12 gpurun_out/c1_fused2_stream.log
