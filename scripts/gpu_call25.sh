#!/bin/bash
# 1-GPU: validate the fp8 tcgen05 GEMM (numerics + timing), then the full gpu suite + smoke on the final tree
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout -k 10 150 python -m pytest tests/test_kernels/test_gemm_fp8.py -x -q > gpurun_out/fp8_test.log 2>&1; echo "fp8 test rc=$?"; tail -4 gpurun_out/fp8_test.log | cut -c1-300
timeout -k 10 90 python scripts/bench_fp8_gemm.py > gpurun_out/fp8_bench.log 2>&1; echo "fp8 bench rc=$?"; grep FP8_GEMM gpurun_out/fp8_bench.log | cut -c1-400; tail -2 gpurun_out/fp8_bench.log | cut -c1-300
timeout -k 10 300 python -m pytest tests -m gpu -x -q --deselect tests/test_kernels/test_gemm_fp8.py > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final2.log | cut -c1-300
timeout -k 10 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final2.log | cut -c1-300
