#!/bin/bash
# BASELINE config 3: Llama-3-70B shape, TP4 x PP2 on 8 GPUs, ZeRO-1 optimizer state tiered to pinned host memory.
# The layer count adapts to the host memory of the box (full 80 layers need ~430 GB of pinned memory at frac 0.5).
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
MEM_GB=$(awk '/MemTotal/ {printf "%d", $2/1048576}' /proc/meminfo)
if [ "$MEM_GB" -ge 1400 ]; then LAYERS=80; else LAYERS=40; fi
echo "host memory ${MEM_GB} GB -> ${LAYERS} layers" > gpurun_out/c12_70b.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 540 $TR --master-port 29671 examples/language/llama/benchmark.py -c llama3-70b --layers $LAYERS -p 3d --tp 4 --pp 2 --zero 1 \
  --offload_optim_frac 0.5 --sp_mode split_gather --comm_backend fused -b 8 --mbs 1 -l 4096 -s 4 -i 2 -g >> gpurun_out/c12_70b.log 2>&1
echo "70b rc=$?" >> gpurun_out/c12_70b.log
grep -E "host memory|model |step |throughput|peak|rc=|Error|error" gpurun_out/c12_70b.log | tail -14 | cut -c1-300
