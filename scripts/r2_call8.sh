#!/bin/bash
# round-2 call 8 (1 GPU): re-run of the flash varlen tests (warp-collective tcgen05.ld fix) and the grouped GEMM backward
# (driver-context fix), grouped GEMM bench, full-model inference bench
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 400 python -m pytest tests/test_kernels/test_flash_attn_native.py -q --timeout 90 > gpurun_out/c8_flash.log 2>&1
echo "flash rc=$?" >> gpurun_out/c8_flash.log
timeout 300 python -m pytest tests/test_kernels/test_grouped_gemm.py -q --timeout 90 > gpurun_out/c8_grouped.log 2>&1
echo "grouped rc=$?" >> gpurun_out/c8_grouped.log
timeout 300 python -m pytest tests/test_kernels/test_inference_kernels.py tests/test_zero/test_zero_offload_gpu.py -q --timeout 120 > gpurun_out/c8_infer_tests.log 2>&1
echo "infer_tests rc=$?" >> gpurun_out/c8_infer_tests.log
timeout 240 python scripts/bench_grouped_gemm.py > gpurun_out/c8_grouped_bench.log 2>&1
timeout 400 python examples/inference/benchmark_llama.py -m llama3-8b -b 16 --in_len 512 --out_len 64 --cuda_graph > gpurun_out/c8_infer_8b.log 2>&1
echo "infer rc=$?" >> gpurun_out/c8_infer_8b.log
timeout 300 python examples/inference/benchmark_llama.py -m llama3-8b -b 64 --in_len 1024 --out_len 64 --cuda_graph > gpurun_out/c8_infer_8b_b64.log 2>&1
tail -6 gpurun_out/c8_flash.log | cut -c1-300; tail -6 gpurun_out/c8_grouped.log | cut -c1-300; tail -5 gpurun_out/c8_infer_tests.log | cut -c1-300; grep GROUPED_GEMM gpurun_out/c8_grouped_bench.log | cut -c1-700
tail -2 gpurun_out/c8_infer_8b.log | cut -c1-900; tail -1 gpurun_out/c8_infer_8b_b64.log | cut -c1-900
