#!/bin/bash
# round-2 call 5 (1 GPU): tail-wave N halving of the CTA-pair GEMM: tests, A/B vs whole tiles / K ranges / cuBLAS, N=1 bench A/B
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 600 python -m pytest tests/test_kernels/test_gemm_tcgen05.py tests/test_kernels/test_gemm_fp8.py -q > gpurun_out/c5_gemm_test.log 2>&1
echo "gemm rc=$?" >> gpurun_out/c5_gemm_test.log
timeout 600 python -m pytest tests/test_kernels/test_flash_attn_native.py -q > gpurun_out/c5_flash.log 2>&1
echo "flash rc=$?" >> gpurun_out/c5_flash.log
timeout 600 python -m pytest tests/test_kernels/test_grouped_gemm.py -q > gpurun_out/c5_grouped.log 2>&1
echo "grouped rc=$?" >> gpurun_out/c5_grouped.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_kernels/test_gemm_tcgen05.py --deselect tests/test_kernels/test_flash_attn_native.py --deselect tests/test_kernels/test_grouped_gemm.py > gpurun_out/c5_pytest_rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/c5_pytest_rest.log
timeout 300 python scripts/bench_grouped_gemm.py > gpurun_out/c5_grouped_bench.log 2>&1
timeout 300 python scripts/bench_gemm_ab.py > gpurun_out/c5_gemm_ab.log 2>&1
CB200_GEMM_BACKEND=native timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e > gpurun_out/c5_bench_native.log 2>&1
CB200_GEMM_BACKEND=cublas timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e > gpurun_out/c5_bench_cublas.log 2>&1
CB200_GEMM_BACKEND=native timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e > gpurun_out/c5_bench_native2.log 2>&1
tail -3 gpurun_out/c5_gemm_test.log; tail -8 gpurun_out/c5_pytest_rest.log; tail -12 gpurun_out/c5_flash.log; tail -12 gpurun_out/c5_grouped.log; cat gpurun_out/c5_grouped_bench.log; cat gpurun_out/c5_gemm_ab.log
python - <<'PY'
import json
for f in ("gpurun_out/c5_bench_native.log","gpurun_out/c5_bench_cublas.log","gpurun_out/c5_bench_native2.log"):
    for l in open(f):
        if l.startswith("{"):
            r=json.loads(l); print(f, round(r["value"]), round(r["ms_per_step"],1), r["clocks"])
PY
