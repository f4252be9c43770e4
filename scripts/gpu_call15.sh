#!/bin/bash
# 1-GPU: main-loop variants of the CTA-pair GEMM (correctness + speed)
export PYTHONPATH=$PWD
mkdir -p gpurun_out
for v in 1 2; do
  CB200_GEMM_2CTA_VARIANT=$v timeout -k 10 300 python -m pytest tests/test_kernels/test_gemm_tcgen05.py -x -q -k "512 or accumulate" > gpurun_out/pytest_gemm_v$v.log 2>&1; echo "variant $v tests rc=$?"; tail -2 gpurun_out/pytest_gemm_v$v.log | cut -c1-200
done
for v in 0 1 2; do
  CB200_GEMM_2CTA_VARIANT=$v timeout -k 10 300 python tests/bench_gemm.py > gpurun_out/gemm_bench_v$v.jsonl 2>/dev/null; echo "variant $v bench rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/gemm_bench_v$v.jsonl"):
    r=json.loads(l); print($v, r["name"], "2cta %.0f"%r["nt_2cta_tflops"], "nn %.0f"%r["nn"]["ours_tflops"], "tn %.0f"%r["tn"]["ours_tflops"], "cublas nt %.0f"%r["nt"]["cublas_tflops"])
PY
done
