#!/bin/bash
# round-2 8-GPU call B, most important first: (1) BASELINE config 3: Llama-3-70B, TP4 x PP2 on 8 GPUs, ZeRO-1 optimizer
# state tiered to pinned host memory (full 80 layers when the box has the host memory for it; NCCL comm backend: the
# fused kernels have not been exercised on TP sub-groups yet), (2) ring attention timing at 16k local tokens with the
# staged forward, (3) Llama-3-8B 128k sp=8 step with the fixed ring forward
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
MEM_GB=$(awk '/MemTotal/ {printf "%d", $2/1048576}' /proc/meminfo)
if [ "$MEM_GB" -ge 900 ]; then LAYERS=80; else LAYERS=40; fi
echo "host memory ${MEM_GB} GB -> ${LAYERS} layers, offload_optim_frac 0.25" > gpurun_out/c12_70b.log
timeout 360 $TR --master-port 29671 examples/language/llama/benchmark.py -c llama3-70b --layers $LAYERS -p 3d --tp 4 --pp 2 --zero 1 \
  --offload_optim_frac 0.25 --sp_mode split_gather --comm_backend nccl -b 8 --mbs 1 -l 4096 -s 3 -i 1 -g >> gpurun_out/c12_70b.log 2>&1
echo "70b rc=$?" >> gpurun_out/c12_70b.log
CB200_RING_ATTN_TIMING=1 CB200_RING_LOCAL_TOKENS=16384 NGPU=8 timeout 150 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c12_ring8.log 2>&1
echo "ring8 rc=$?" >> gpurun_out/c12_ring8.log
CB200_RING_ATTN=fused timeout 200 $TR --master-port 29651 examples/language/llama/benchmark.py -c llama3-8b -p 3d --sp 8 --sp_mode ring_attn --zero 1 \
  -b 1 -l 131072 -s 3 -i 1 -g > gpurun_out/c12_llama128k_fused.log 2>&1
echo "llama128k_fused rc=$?" >> gpurun_out/c12_llama128k_fused.log
grep -E "host memory|model |step |throughput|peak|rc=|Error|error" gpurun_out/c12_70b.log | tail -12 | cut -c1-300
grep -E "RING_TIMING|RING_ATTN_GPU_OK|rc=|Error|Mismatch" gpurun_out/c12_ring8.log | cut -c1-600 | tail -4
grep -E "step |throughput|peak|rc=|Error" gpurun_out/c12_llama128k_fused.log | tail -6 | cut -c1-300
