#!/bin/bash
# round-2 call 9 (8 GPUs): BASELINE configs 5 and 4 - ring attention at 128k context (sp=8) and expert parallel ep=8
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
# (a) one attention layer, 16k local tokens per rank: fused ring vs python ring (P2P gather) vs python ring (NCCL)
CB200_RING_ATTN_TIMING=1 CB200_RING_LOCAL_TOKENS=16384 NGPU=8 timeout 420 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c9_ring8.log 2>&1
echo "ring8 rc=$?" >> gpurun_out/c9_ring8.log
# (b) full Llama-3-8B training step at 131072 tokens, sp=8 ring attention + ZeRO-1 over the dp x sp group
CB200_RING_ATTN=fused timeout 420 $TR --master-port 29651 examples/language/llama/benchmark.py -c llama3-8b -p 3d --sp 8 --sp_mode ring_attn --zero 1 \
  -b 1 -l 131072 -s 5 -i 2 > gpurun_out/c9_llama128k_fused.log 2>&1
echo "llama128k_fused rc=$?" >> gpurun_out/c9_llama128k_fused.log
if ! grep -q "throughput" gpurun_out/c9_llama128k_fused.log; then
  CB200_RING_ATTN=fused timeout 420 $TR --master-port 29652 examples/language/llama/benchmark.py -c llama3-8b -p 3d --sp 8 --sp_mode ring_attn --zero 1 \
    -b 1 -l 131072 -s 5 -i 2 -g > gpurun_out/c9_llama128k_fused_ckpt.log 2>&1
  echo "llama128k_fused_ckpt rc=$?" >> gpurun_out/c9_llama128k_fused_ckpt.log
fi
CB200_RING_ATTN=python timeout 420 $TR --master-port 29653 examples/language/llama/benchmark.py -c llama3-8b -p 3d --sp 8 --sp_mode ring_attn --zero 1 \
  -b 1 -l 131072 -s 4 -i 2 > gpurun_out/c9_llama128k_python.log 2>&1
echo "llama128k_python rc=$?" >> gpurun_out/c9_llama128k_python.log
# (c) expert parallel ep=8: MoE layer micro-benchmark and a 4-layer Mixtral-8x7B-width training step
timeout 420 $TR --master-port 29654 scripts/bench_moe_ep.py > gpurun_out/c9_moe_ep8.log 2>&1
echo "moe_ep8 rc=$?" >> gpurun_out/c9_moe_ep8.log
timeout 420 $TR --master-port 29655 examples/language/mixtral/benchmark.py -c mixtral-8x7b --layers 4 --ep 8 -b 2 -l 4096 > gpurun_out/c9_mixtral_fused.log 2>&1
timeout 420 $TR --master-port 29656 examples/language/mixtral/benchmark.py -c mixtral-8x7b --layers 4 --ep 8 -b 2 -l 4096 --moe_backend nccl --grouped_gemm lib > gpurun_out/c9_mixtral_nccl.log 2>&1
grep -E "RING_TIMING|RING_ATTN_GPU_OK|rc=|Error" gpurun_out/c9_ring8.log | cut -c1-700 | tail -5
for f in gpurun_out/c9_llama128k_*.log; do echo $f; grep -E "throughput|rc=|Error|OutOfMemory" $f | tail -3 | cut -c1-300; done
grep -E "MOE_EP|rc=|Error" gpurun_out/c9_moe_ep8.log | cut -c1-1000; grep -E "throughput|MOE_BENCH|Error" gpurun_out/c9_mixtral_*.log | cut -c1-600
