#!/bin/bash
# 2-GPU: tile-granular GEMM->RS correctness + timing; N=2 bench + profile; N=1 bench; ncu of the 2-CTA GEMM
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=2 timeout -k 10 420 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused2c.log 2>&1; echo "fused rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch" gpurun_out/fused2c.log | cut -c1-600 | tail -12
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --profile gpurun_out/prof_n2_c.txt > gpurun_out/b2d_fused.log 2>&1; echo "bench2 rc=$?"; grep -a '"metric"' gpurun_out/b2d_fused.log | cut -c1-500
head -12 gpurun_out/prof_n2_c.txt | cut -c1-150
CUDA_VISIBLE_DEVICES=0 timeout -k 10 900 python bench.py --steps 3 --warmup 3 --no-e2e --profile gpurun_out/prof_n1_c.txt > gpurun_out/b1_c.log 2>&1; echo "bench1 rc=$?"; grep -a '"metric"' gpurun_out/b1_c.log | cut -c1-500
head -14 gpurun_out/prof_n1_c.txt | cut -c1-150
CUDA_VISIBLE_DEVICES=0 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 4 -c 1 -f -o gpurun_out/gemm_2cta python tests/bench_gemm.py > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu rc=$?"
