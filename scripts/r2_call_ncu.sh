#!/bin/bash
# ncu --set full captures of one launch per kernel family (1 GPU; never a bench number).  The .ncu-rep of the two
# contraction kernels is kept; every capture is also exported as raw / details CSV on the box (small text files).
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
cap() { name=$1; regex=$2; tgt=$3; keep=$4
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -f -o gpurun_out/ncu_$name python scripts/ncu_targets.py $tgt > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"
  if [ -f gpurun_out/ncu_$name.ncu-rep ]; then
    ncu -i gpurun_out/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null
    ncu -i gpurun_out/ncu_$name.ncu-rep --page details --csv > gpurun_out/ncu_${name}_details.csv 2>/dev/null
    [ "$keep" = keep ] || rm -f gpurun_out/ncu_$name.ncu-rep
  fi; }
cap gemm_2cta gemm_tcgen05_2cta_kernel gemm keep
cap flash_fwd flash_fwd_kernel flash_fwd keep
cap flash_bwd flash_bwd_kernel flash_bwd
cap grouped_gemm grouped_gemm_2cta_kernel grouped
cap paged_decode paged_decode_kernel decode
cap rmsnorm rmsnorm_fwd_kernel norm_glu
cap glu glu_fwd_kernel norm_glu
cap adam multi_tensor_adam_kernel adam
cap moe_push ep_push_rows_kernel moe
cap moe_combine ep_pull_combine_kernel moe
cap ce_grad ce_softmax_grad_kernel ce_rope
cap kv_write kv_cache_write_vec_kernel kv_write
ls -la gpurun_out/ncu_* | awk '{print $5, $9}'
