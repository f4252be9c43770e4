#!/bin/bash
# round-2 call 23 (2 GPUs, last GPU seconds): in-place HF path under TP=2 on NCCL
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 50 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 examples/language/hf_inplace/finetune_hf.py --family llama --tp 2 --steps 3 > gpurun_out/c23_hf_llama_tp2.log 2>&1
echo "rc=$?" >> gpurun_out/c23_hf_llama_tp2.log
grep -E "step|rc=|Error|error" gpurun_out/c23_hf_llama_tp2.log | tail -6
