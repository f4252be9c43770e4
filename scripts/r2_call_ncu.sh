#!/bin/bash
# ncu --set full captures of one launch per kernel family (1 GPU; never a bench number)
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
cap() { name=$1; regex=$2; tgt=$3; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -f -o gpurun_out/ncu_$name python scripts/ncu_targets.py $tgt > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"; }
cap gemm_2cta gemm_tcgen05_2cta_kernel gemm
cap flash_fwd flash_fwd_kernel flash_fwd
cap flash_bwd flash_bwd_kernel flash_bwd
cap grouped_gemm grouped_gemm_2cta_kernel grouped
cap rmsnorm rmsnorm_fwd_kernel norm_glu
cap glu glu_fwd_kernel norm_glu
cap adam multi_tensor_adam_kernel adam
cap paged_decode paged_decode_kernel decode
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
