#!/bin/bash
# 2-GPU functional smoke of every plugin / schedule on real GPUs (NCCL, CUDA streams, pinned memory paths)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
B="examples/language/llama/benchmark.py -c llama-100m -b 2 -l 512 -s 3 -i 1"
run() { name=$1; shift; timeout -k 10 240 "$@" > gpurun_out/smoke_$name.log 2>&1; rc=$?; echo "$name rc=$rc $(grep -a 'throughput' gpurun_out/smoke_$name.log | tail -1)"; [ $rc -ne 0 ] && tail -5 gpurun_out/smoke_$name.log | cut -c1-300; }
run ddp     $TR $B -p ddp
run zero2   $TR $B -p zero2
run gemini  $TR $B -p gemini
run fsdp    $TR $B -p fsdp
run tp2sp   $TR $B -p 3d --tp 2 --sp_mode split_gather --comm_backend fused
run ulysses $TR $B -p 3d --sp 2 --sp_mode all_to_all
run ringattn $TR $B -p 3d --sp 2 --sp_mode ring_attn
run pp2     $TR $B -p 3d --pp 2 --mbs 1
run pp2zbv  $TR $B -p 3d --pp 2 --mbs 1 --pp_style zbv --n_chunks 2
run pp2int  $TR $B -p 3d --pp 2 --mbs 1 --pp_style interleaved --n_chunks 2
run moe_ep2 $TR examples/language/mixtral/train.py --ep 2 --zero 1 -s 3
run infer   python examples/inference/benchmark_llama.py -m llama3-8b --layers 4 -b 16 --in_len 256 --out_len 32 --cuda_graph
