#!/bin/bash
# 1-GPU: gpu tests, reference arm, ncu capture of the GEMM kernel
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_n1.log 2>&1; echo "ref rc=$?"
tail -3 gpurun_out/bench_ref_n1.log | cut -c1-1500
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 8 -c 1 -f -o gpurun_out/gemm_tcgen05 python tests/bench_gemm.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
