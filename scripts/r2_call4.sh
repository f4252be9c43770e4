#!/bin/bash
# round-2 call 4 (1 GPU): flash backward first run, GEMM tail-split A/B, GEMM tests
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 600 python -m pytest tests/test_kernels/test_flash_attn_native.py -x -q > gpurun_out/c4_flash.log 2>&1
echo "flash rc=$?" >> gpurun_out/c4_flash.log
timeout 300 python scripts/bench_flash_attn.py > gpurun_out/c4_flash_bench.log 2>&1
timeout 600 python -m pytest tests/test_kernels/test_gemm_tcgen05.py -x -q > gpurun_out/c4_gemm_test.log 2>&1
echo "gemm rc=$?" >> gpurun_out/c4_gemm_test.log
timeout 300 python scripts/bench_gemm_ab.py > gpurun_out/c4_gemm_ab.log 2>&1
tail -15 gpurun_out/c4_flash.log; cat gpurun_out/c4_flash_bench.log | tail -8; tail -3 gpurun_out/c4_gemm_test.log; cat gpurun_out/c4_gemm_ab.log
