#!/bin/bash
# 2-GPU: ticketed staggered RS (all warps), N=2 bench, plugin smoke
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=2 timeout -k 10 300 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused2f.log 2>&1; echo "fused2 rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch\|timeout" gpurun_out/fused2f.log | cut -c1-700 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > gpurun_out/b2g_fused.log 2>&1; echo "bench2 rc=$?"; grep -a '"metric"' gpurun_out/b2g_fused.log | cut -c1-330
bash scripts/gpu_smoke_plugins.sh
grep -a "decode_tokens" gpurun_out/smoke_infer.log | cut -c1-500
