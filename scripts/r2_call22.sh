#!/bin/bash
# round-2 call 22 (1 GPU, last GPU seconds): in-place HF path on the GPU (fused RMSNorm method replacement, bf16)
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29731 examples/language/hf_inplace/finetune_hf.py --family llama --steps 3 > gpurun_out/c22_hf_llama.log 2>&1
echo "rc=$?" >> gpurun_out/c22_hf_llama.log
grep -E "step|rc=|Error|error" gpurun_out/c22_hf_llama.log | tail -6
