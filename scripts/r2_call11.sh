#!/bin/bash
# round-2 call 11 (2 GPUs): validate the aligned fused-EP layout, re-time MoE layer + grouped GEMM pieces, Gemini after
# the host-sync removal, decode-step kernel breakdown
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 300 python -m pytest tests/test_moe tests/test_kernels/test_grouped_gemm.py -m gpu -q -x --timeout 120 > gpurun_out/c11_moe_tests.log 2>&1
echo "moe_tests rc=$?" >> gpurun_out/c11_moe_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29631 scripts/bench_moe_ep.py > gpurun_out/c11_moe_ep2.log 2>&1
echo "moe_ep2 rc=$?" >> gpurun_out/c11_moe_ep2.log
timeout 300 python scripts/bench_grouped_gemm.py > gpurun_out/c11_grouped_bench.log 2>&1
B="examples/language/llama/benchmark.py -b 4 -l 1024 -s 10 -i 4"
for spec in "zero2_100m:-c llama-100m -p zero2" "gemini_100m:-c llama-100m -p gemini" "zero2_5b:-c llama-5b -p zero2" "gemini_5b:-c llama-5b -p gemini"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout -k 10 240 $TR --master-port 29641 $B $args > gpurun_out/c11_smoke_$name.log 2>&1
  echo "$name rc=$? $(grep -a 'throughput' gpurun_out/c11_smoke_$name.log | tail -1)" >> gpurun_out/c11_plugins.log
done
timeout 400 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 32 --cuda_graph --profile gpurun_out/c11_decode_breakdown_b16.txt > gpurun_out/c11_infer_b16.log 2>&1
timeout 400 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 32 --cuda_graph --profile gpurun_out/c11_decode_breakdown_b64.txt > gpurun_out/c11_infer_b64.log 2>&1
tail -5 gpurun_out/c11_moe_tests.log; grep -E "MOE_EP|rc=|Error" gpurun_out/c11_moe_ep2.log | cut -c1-900
grep GROUPED gpurun_out/c11_grouped_bench.log | cut -c1-1200; cat gpurun_out/c11_plugins.log
head -14 gpurun_out/c11_decode_breakdown_b16.txt; head -10 gpurun_out/c11_decode_breakdown_b64.txt; tail -2 gpurun_out/c11_infer_b16.log | cut -c1-600
