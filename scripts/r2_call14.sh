#!/bin/bash
# round-2 call 14 (2 GPUs): isolate the ring-attention forward mismatch (block-mode kernel test on one GPU; ring test with
# and without the staged forward), decode-kernel pipeline (tests + Llama-3-8B bench), MoE padded NCCL path
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 300 python -m pytest tests/test_kernels/test_flash_attn_native.py -m gpu -q --timeout 120 -k "block_mode" > gpurun_out/c14_block.log 2>&1
echo "block rc=$?" >> gpurun_out/c14_block.log
CB200_RING_STAGE=0 NGPU=2 timeout 200 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c14_ring_direct.log 2>&1
echo "ring_direct rc=$?" >> gpurun_out/c14_ring_direct.log
CB200_RING_ATTN_TIMING=1 CB200_RING_LOCAL_TOKENS=8192 NGPU=2 timeout 300 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c14_ring_staged.log 2>&1
echo "ring_staged rc=$?" >> gpurun_out/c14_ring_staged.log
timeout 300 python -m pytest tests/test_kernels/test_inference_kernels.py tests/test_moe -m gpu -q -x --timeout 200 > gpurun_out/c14_infer_moe_tests.log 2>&1
echo "infer_moe rc=$?" >> gpurun_out/c14_infer_moe_tests.log
timeout 300 python examples/inference/benchmark_llama.py -b 16 --in_len 512 --out_len 32 --cuda_graph --profile gpurun_out/c14_decode_breakdown_b16.txt > gpurun_out/c14_infer_b16.log 2>&1
timeout 300 python examples/inference/benchmark_llama.py -b 64 --in_len 1024 --out_len 32 --cuda_graph --profile gpurun_out/c14_decode_breakdown_b64.txt > gpurun_out/c14_infer_b64.log 2>&1
grep -E "passed|failed|Error|rc=|Mismatch|Greatest" gpurun_out/c14_block.log | tail -12
for f in direct staged; do echo "== ring $f"; grep -E "RING_TIMING|RING_ATTN_GPU_OK|rc=|Mismatched|Greatest|Error" gpurun_out/c14_ring_$f.log | cut -c1-500 | tail -6; done
tail -4 gpurun_out/c14_infer_moe_tests.log; grep -h '"model"' gpurun_out/c14_infer_b16.log gpurun_out/c14_infer_b64.log | cut -c1-600
head -4 gpurun_out/c14_decode_breakdown_b16.txt; head -3 gpurun_out/c14_decode_breakdown_b64.txt
