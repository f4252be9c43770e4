#!/bin/bash
# 1-GPU: 2-CTA GEMM correctness + speed, optimizer norm kernel, MoE/softmax/fp8/inference gpu tests, N=1 bench w/ accum
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels/test_gemm_tcgen05.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -6 gpurun_out/pytest_gemm.log | cut -c1-300
timeout -k 10 300 python tests/bench_gemm.py > gpurun_out/gemm_bench2.jsonl 2> gpurun_out/gemm_bench2.err; echo "bench_gemm rc=$?"; cut -c1-900 gpurun_out/gemm_bench2.jsonl
timeout -k 10 900 python -m pytest tests -m gpu -q --deselect tests/test_kernels/test_gemm_tcgen05.py > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu3.log | cut -c1-300
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --profile gpurun_out/prof_n1_accum.txt > gpurun_out/b1_accum.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/b1_accum.log | cut -c1-1500
head -12 gpurun_out/prof_n1_accum.txt | cut -c1-160
