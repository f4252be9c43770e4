#!/bin/bash
# 2-GPU: staggered pull-accumulate RS (correctness+timing), GEMM main-loop variants, plugin smoke, N=2 bench
export PYTHONPATH=$PWD
mkdir -p gpurun_out
NGPU=2 timeout -k 10 300 python tests/test_parallel/test_fused_comm.py > gpurun_out/fused2d.log 2>&1; echo "fused2 rc=$?"; grep -a "FUSED_\|Error\|error\|Mismatch\|timeout" gpurun_out/fused2d.log | cut -c1-700 | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout -k 10 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --profile gpurun_out/prof_n2_d.txt > gpurun_out/b2e_fused.log 2>&1; echo "bench2 rc=$?"; grep -a '"metric"' gpurun_out/b2e_fused.log | cut -c1-400
head -8 gpurun_out/prof_n2_d.txt | cut -c1-150
for v in 0 1 2; do
  CUDA_VISIBLE_DEVICES=0 CB200_GEMM_2CTA_VARIANT=$v timeout -k 10 200 python tests/bench_gemm.py > gpurun_out/gemm_bench_v$v.jsonl 2>/dev/null; echo "variant $v bench rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/gemm_bench_v$v.jsonl"):
    r=json.loads(l); print($v, r["name"], "2cta %.0f"%r["nt_2cta_tflops"], "nn %.0f"%r["nn"]["ours_tflops"], "tn %.0f"%r["tn"]["ours_tflops"], "cublas nt %.0f"%r["nt"]["cublas_tflops"])
PY
done
CUDA_VISIBLE_DEVICES=0 CB200_GEMM_2CTA_VARIANT=2 timeout -k 10 200 python -m pytest tests/test_kernels/test_gemm_tcgen05.py -x -q -k "512 or accumulate" > gpurun_out/pytest_gemm_v2.log 2>&1; echo "variant2 tests rc=$?"; tail -2 gpurun_out/pytest_gemm_v2.log | cut -c1-200
bash scripts/gpu_smoke_plugins.sh
