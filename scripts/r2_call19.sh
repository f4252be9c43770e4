#!/bin/bash
# round-2 call 19 (1 GPU): tensor-core decode kernel as the default - kernel grid both ways, engine tests, decode bench
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 200 python -m pytest tests/test_kernels/test_inference_kernels.py tests/test_infer_engine.py -m gpu -q --timeout 100 > gpurun_out/c19_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c19_tests.log
CB200_DECODE=simt timeout 100 python -m pytest tests/test_kernels/test_inference_kernels.py -m gpu -q --timeout 100 -k decode > gpurun_out/c19_tests_simt.log 2>&1
echo "tests_simt rc=$?" >> gpurun_out/c19_tests_simt.log
timeout 150 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 64 --cuda_graph > gpurun_out/c19_infer_b64.log 2>&1
timeout 150 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 64 --cuda_graph > gpurun_out/c19_infer_b16.log 2>&1
tail -3 gpurun_out/c19_tests.log | cut -c1-200; tail -3 gpurun_out/c19_tests_simt.log | cut -c1-200
grep -h '"model"' gpurun_out/c19_infer_b64.log gpurun_out/c19_infer_b16.log | cut -c1-700
