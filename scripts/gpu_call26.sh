#!/bin/bash
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout -k 10 110 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta --launch-skip 3 --launch-count 1 -o gpurun_out/gemm_fp8_2cta -f python scripts/ncu_fp8_gemm.py > gpurun_out/ncu_fp8.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fp8.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep 2>/dev/null
