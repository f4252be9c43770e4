#!/bin/bash
# round-2 call 16 (1 GPU): decode kernel with warps = kv heads (tests + Llama-3-8B bench + ncu of the kernel)
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 300 python -m pytest tests/test_kernels/test_inference_kernels.py tests/test_infer_engine.py -m gpu -q -x --timeout 120 > gpurun_out/c16_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c16_tests.log
timeout 300 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 32 --cuda_graph --profile gpurun_out/c16_decode_breakdown_b16.txt > gpurun_out/c16_infer_b16.log 2>&1
timeout 300 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 32 --cuda_graph --profile gpurun_out/c16_decode_breakdown_b64.txt > gpurun_out/c16_infer_b64.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:paged_decode_kernel -s 2 -c 1 -f -o gpurun_out/ncu_paged_decode_v3 python scripts/ncu_targets.py decode > gpurun_out/c16_ncu.log 2>&1
ncu -i gpurun_out/ncu_paged_decode_v3.ncu-rep --page raw --csv > gpurun_out/ncu_paged_decode_v3_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_paged_decode_v3.ncu-rep --page details --csv > gpurun_out/ncu_paged_decode_v3_details.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:flash_fwd_kernel -s 2 -c 1 -f -o gpurun_out/ncu_flash_fwd_v2 python scripts/ncu_targets.py flash_fwd > gpurun_out/c16_ncu2.log 2>&1
ncu -i gpurun_out/ncu_flash_fwd_v2.ncu-rep --page raw --csv > gpurun_out/ncu_flash_fwd_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_flash_fwd_v2.ncu-rep --page details --csv > gpurun_out/ncu_flash_fwd_v2_details.csv 2>/dev/null
tail -4 gpurun_out/c16_tests.log; grep -h '"model"' gpurun_out/c16_infer_b16.log gpurun_out/c16_infer_b64.log | cut -c1-600
head -4 gpurun_out/c16_decode_breakdown_b16.txt; head -3 gpurun_out/c16_decode_breakdown_b64.txt
