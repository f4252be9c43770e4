#!/bin/bash
# round-2 call 13 (2 GPUs): flash forward rewrite (accumulator in TMEM) - numerics + timing vs cuDNN; ring attention with
# the staged forward (tests + timing); MoE tests (padded layout on the NCCL path); ring-attn / fused-comm model oracles
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 400 python -m pytest tests/test_kernels/test_flash_attn_native.py -m gpu -q -x --timeout 120 > gpurun_out/c13_flash_tests.log 2>&1
echo "flash_tests rc=$?" >> gpurun_out/c13_flash_tests.log
timeout 300 python scripts/bench_flash_attn.py > gpurun_out/c13_flash_bench.log 2>&1
echo "flash_bench rc=$?" >> gpurun_out/c13_flash_bench.log
CB200_RING_ATTN_TIMING=1 CB200_RING_LOCAL_TOKENS=8192 NGPU=2 timeout 300 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c13_ring2.log 2>&1
echo "ring2 rc=$?" >> gpurun_out/c13_ring2.log
timeout 400 python -m pytest tests/test_moe tests/test_shardformer/test_ring_attention.py -m gpu -q -x --timeout 200 > gpurun_out/c13_moe_ring_tests.log 2>&1
echo "moe_ring_tests rc=$?" >> gpurun_out/c13_moe_ring_tests.log
tail -4 gpurun_out/c13_flash_tests.log; grep -E "FLASH|rc=|Error" gpurun_out/c13_flash_bench.log | cut -c1-500 | tail -12
grep -E "RING_TIMING|RING_ATTN_GPU_OK|rc=|Error" gpurun_out/c13_ring2.log | cut -c1-600 | tail -5; tail -5 gpurun_out/c13_moe_ring_tests.log
