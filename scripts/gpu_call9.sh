#!/bin/bash
# 2-GPU: per-kernel profile of one step (fused + nccl backends) + MoE fused EP test + FSDP test
export PYTHONPATH=$PWD
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --profile gpurun_out/prof_n2_fused.txt > gpurun_out/b2p_fused.log 2>&1; echo "fused rc=$?"; tail -1 gpurun_out/b2p_fused.log | cut -c1-300
timeout -k 10 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --comm-backend nccl --profile gpurun_out/prof_n2_nccl.txt > gpurun_out/b2p_nccl.log 2>&1; echo "nccl rc=$?"; tail -1 gpurun_out/b2p_nccl.log | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-e2e --profile gpurun_out/prof_n1.txt > gpurun_out/b1p.log 2>&1; echo "n1 rc=$?"; tail -1 gpurun_out/b1p.log | cut -c1-300
NGPU=2 timeout -k 10 300 python tests/test_moe/test_moe_ops.py > gpurun_out/moe_ep2.log 2>&1; echo "moe rc=$?"; tail -5 gpurun_out/moe_ep2.log
timeout -k 10 600 python -m pytest tests/test_moe tests/test_booster/test_dp_plugins.py tests/test_fp8 tests/test_kernels/test_inference_kernels.py -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu2.log
head -45 gpurun_out/prof_n2_fused.txt
