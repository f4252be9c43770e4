#!/bin/bash
# round-2 evidence call (1 GPU): ncu --set full per kernel family + compute-sanitizer memcheck / racecheck of kernel tests
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
bash scripts/r2_call_ncu.sh > gpurun_out/ev_ncu.log 2>&1
san() { tool=$1; name=$2; shift 2; timeout 240 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 10 python -m pytest "$@" -q -m gpu -x > gpurun_out/sanitize_${tool}_${name}.log 2>&1; echo "sanitize $tool $name rc=$?" >> gpurun_out/ev_sanitize.log; }
: > gpurun_out/ev_sanitize.log
san memcheck norm tests/test_kernels/test_elementwise_norm.py
san memcheck gemm tests/test_kernels/test_gemm_tcgen05.py -k "test_gemm_nt and 128-256-64 and bfloat16"
san memcheck flash tests/test_kernels/test_flash_attn_native.py -k "(fwd_matches and 1-128-2-2-128) or (bwd_matches and 1-128-2-2) or (varlen and lens2)"
san memcheck grouped tests/test_kernels/test_grouped_gemm.py -k "counts0"
san memcheck infer tests/test_kernels/test_inference_kernels.py -k "sliding_window or kv_cache_write"
san racecheck norm tests/test_kernels/test_elementwise_norm.py
san racecheck softmax tests/test_kernels/test_softmax_kernels.py
san racecheck optim tests/test_kernels/test_optim_kernels.py
cat gpurun_out/ev_ncu.log | tail -12; cat gpurun_out/ev_sanitize.log
for f in gpurun_out/sanitize_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $f | tail -3; done
