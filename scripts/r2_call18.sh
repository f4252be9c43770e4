#!/bin/bash
# round-2 call 18 (1 GPU): tensor-core decode kernel (opt-in) - numerics grid + Llama-3-8B decode vs the default kernel
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
CB200_DECODE=mma timeout 200 python -m pytest tests/test_kernels/test_inference_kernels.py -m gpu -q --timeout 100 > gpurun_out/c18_tests_mma.log 2>&1
echo "tests_mma rc=$?" >> gpurun_out/c18_tests_mma.log
CB200_DECODE=mma timeout 150 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 32 --cuda_graph --profile gpurun_out/c18_decode_breakdown_b64_mma.txt > gpurun_out/c18_infer_b64_mma.log 2>&1
timeout 150 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 32 --cuda_graph --profile gpurun_out/c18_decode_breakdown_b64.txt > gpurun_out/c18_infer_b64.log 2>&1
CB200_DECODE=mma timeout 150 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 32 --cuda_graph --profile gpurun_out/c18_decode_breakdown_b16_mma.txt > gpurun_out/c18_infer_b16_mma.log 2>&1
timeout 150 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 32 --cuda_graph --profile gpurun_out/c18_decode_breakdown_b16.txt > gpurun_out/c18_infer_b16.log 2>&1
tail -6 gpurun_out/c18_tests_mma.log | cut -c1-300
for f in b64_mma b64 b16_mma b16; do echo "== $f"; grep -h '"model"' gpurun_out/c18_infer_$f.log | cut -c1-420; grep paged_decode gpurun_out/c18_decode_breakdown_$f.txt | cut -c1-120; done
