#!/bin/bash
# round-2 combined 8-GPU call, most important first: (1) N=8 bench (tp8+sp row, zero1(dp8) row) with kernel breakdown,
# (2) TP=8 fused comm kernels vs NCCL+cuBLAS, (3) expert parallel ep=8, (4) ring attention timing at 16k local tokens,
# (5) Llama-3-8B 128k sp=8 step, (6) reference arm smoke at N=8, (7) Mixtral benchmark
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 --profile gpurun_out/c10_prof_n8_tp.txt > gpurun_out/c10_bench_n8.log 2>&1
echo "bench8 rc=$?" >> gpurun_out/c10_bench_n8.log
NGPU=8 timeout 300 python tests/test_parallel/test_fused_comm.py > gpurun_out/c10_fused8.log 2>&1
echo "fused8 rc=$?" >> gpurun_out/c10_fused8.log
timeout 240 $TR --master-port 29654 scripts/bench_moe_ep.py > gpurun_out/c10_moe_ep8.log 2>&1
echo "moe_ep8 rc=$?" >> gpurun_out/c10_moe_ep8.log
CB200_RING_ATTN_TIMING=1 CB200_RING_LOCAL_TOKENS=16384 NGPU=8 timeout 240 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c10_ring8.log 2>&1
echo "ring8 rc=$?" >> gpurun_out/c10_ring8.log
CB200_RING_ATTN=fused timeout 300 $TR --master-port 29651 examples/language/llama/benchmark.py -c llama3-8b -p 3d --sp 8 --sp_mode ring_attn --zero 1 \
  -b 1 -l 131072 -s 4 -i 2 -g > gpurun_out/c10_llama128k_fused.log 2>&1
echo "llama128k_fused rc=$?" >> gpurun_out/c10_llama128k_fused.log
timeout 400 $TR --master-port 29657 bench.py --impl reference --gpus 8 --steps 2 --warmup 3 --layers 4 --parallelism tp > gpurun_out/c10_bench_ref8.log 2>&1
echo "bench_ref8 rc=$?" >> gpurun_out/c10_bench_ref8.log
timeout 240 $TR --master-port 29655 examples/language/mixtral/benchmark.py -c mixtral-8x7b --layers 4 --ep 8 -b 2 -l 4096 > gpurun_out/c10_mixtral_fused.log 2>&1
tail -c 3000 gpurun_out/c10_bench_n8.log
grep -E "FUSED_TIMING|RS_TUNING|FUSED_STATS|rc=|Error|timeout" gpurun_out/c10_fused8.log | cut -c1-700 | tail -22
grep -E "MOE_EP|rc=|Error" gpurun_out/c10_moe_ep8.log | cut -c1-1000
grep -E "RING_TIMING|RING_ATTN_GPU_OK|rc=|Error" gpurun_out/c10_ring8.log | cut -c1-700 | tail -5
for f in gpurun_out/c10_llama128k_*.log; do echo $f; grep -E "throughput|rc=|Error|OutOfMemory" $f | tail -3 | cut -c1-300; done
tail -c 1200 gpurun_out/c10_bench_ref8.log; grep -E "throughput|MOE_BENCH|Error" gpurun_out/c10_mixtral_*.log | cut -c1-600
