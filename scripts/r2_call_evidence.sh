#!/bin/bash
# round-2 evidence call (1 GPU): (a) GPU tests of the kernels changed after the last full run, (b) Llama-3-8B inference
# bench after the decode-kernel rewrite, (c) ncu --set full per kernel family, (d) compute-sanitizer memcheck / racecheck
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 420 python -m pytest tests/test_kernels/test_inference_kernels.py tests/test_kernels/test_flash_attn_native.py tests/test_moe tests/test_kernels/test_grouped_gemm.py -m gpu -q -x --timeout 120 > gpurun_out/ev_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ev_tests.log
timeout 300 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 32 --cuda_graph --profile gpurun_out/ev_decode_breakdown_b16.txt > gpurun_out/ev_infer_b16.log 2>&1
timeout 300 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 32 --cuda_graph --profile gpurun_out/ev_decode_breakdown_b64.txt > gpurun_out/ev_infer_b64.log 2>&1
bash scripts/r2_call_ncu.sh > gpurun_out/ev_ncu.log 2>&1
san() { tool=$1; name=$2; shift 2; timeout 200 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 10 python -m pytest "$@" -q -m gpu -x > gpurun_out/sanitize_${tool}_${name}.log 2>&1; echo "sanitize $tool $name rc=$?" >> gpurun_out/ev_sanitize.log; }
: > gpurun_out/ev_sanitize.log
san memcheck norm tests/test_kernels/test_elementwise_norm.py
san memcheck gemm tests/test_kernels/test_gemm_tcgen05.py -k "test_gemm_nt and 128-256-64 and bfloat16"
san memcheck flash tests/test_kernels/test_flash_attn_native.py -k "(fwd_matches and 1-128-2-2-128) or (bwd_matches and 1-128-2-2) or (varlen and lens2 and True and 4-4)"
san memcheck grouped tests/test_kernels/test_grouped_gemm.py -k "counts0"
san memcheck infer tests/test_kernels/test_inference_kernels.py -k "sliding_window or kv_cache_write"
san racecheck norm tests/test_kernels/test_elementwise_norm.py
san racecheck softmax tests/test_kernels/test_softmax_kernels.py
san racecheck optim tests/test_kernels/test_optim_kernels.py
tail -4 gpurun_out/ev_tests.log; grep -h '"model"' gpurun_out/ev_infer_b16.log gpurun_out/ev_infer_b64.log | cut -c1-700
head -6 gpurun_out/ev_decode_breakdown_b16.txt; head -5 gpurun_out/ev_decode_breakdown_b64.txt
tail -16 gpurun_out/ev_ncu.log; cat gpurun_out/ev_sanitize.log
for f in gpurun_out/sanitize_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|error" $f | tail -3; done
