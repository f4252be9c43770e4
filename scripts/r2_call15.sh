#!/bin/bash
# round-2 call 15 (2 GPUs): flash forward with the accumulator-barrier fix (flash tests, ring attention tests + timing),
# decode kernel on the cp.async ring (tests + Llama-3-8B bench)
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 400 python -m pytest tests/test_kernels/test_flash_attn_native.py tests/test_kernels/test_inference_kernels.py -m gpu -q -x --timeout 120 > gpurun_out/c15_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c15_tests.log
CB200_RING_ATTN_TIMING=1 CB200_RING_LOCAL_TOKENS=8192 NGPU=2 timeout 300 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c15_ring_staged.log 2>&1
echo "ring_staged rc=$?" >> gpurun_out/c15_ring_staged.log
CB200_RING_STAGE=0 NGPU=2 timeout 200 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c15_ring_direct.log 2>&1
echo "ring_direct rc=$?" >> gpurun_out/c15_ring_direct.log
timeout 300 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 32 --cuda_graph --profile gpurun_out/c15_decode_breakdown_b16.txt > gpurun_out/c15_infer_b16.log 2>&1
timeout 300 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 32 --cuda_graph --profile gpurun_out/c15_decode_breakdown_b64.txt > gpurun_out/c15_infer_b64.log 2>&1
timeout 200 python scripts/bench_flash_attn.py > gpurun_out/c15_flash_bench.log 2>&1
tail -4 gpurun_out/c15_tests.log
for f in staged direct; do echo "== ring $f"; grep -E "RING_TIMING|RING_ATTN_GPU_OK|rc=|Mismatched|Greatest|Error" gpurun_out/c15_ring_$f.log | cut -c1-500 | tail -6; done
grep -h '"model"' gpurun_out/c15_infer_b16.log gpurun_out/c15_infer_b64.log | cut -c1-600
head -4 gpurun_out/c15_decode_breakdown_b16.txt; head -3 gpurun_out/c15_decode_breakdown_b64.txt; grep FLASH_FWD gpurun_out/c15_flash_bench.log | cut -c1-300
