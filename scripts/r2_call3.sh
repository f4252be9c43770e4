#!/bin/bash
# round-2 call 3 (1 GPU): full GPU test suite, GEMM tail-split micro-benchmark at M=4096, N=1 bench native vs cuBLAS GEMM
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c3_pytest_gpu.log
CB200_GEMM_BENCH_M=4096 timeout 300 python tests/bench_gemm.py > gpurun_out/c3_gemm_m4096_split.jsonl 2> gpurun_out/c3_gemm.err
CB200_GEMM_TAIL_SPLIT=0 CB200_GEMM_BENCH_M=4096 timeout 300 python tests/bench_gemm.py > gpurun_out/c3_gemm_m4096_nosplit.jsonl 2>> gpurun_out/c3_gemm.err
CB200_GEMM_BACKEND=native timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e --profile gpurun_out/c3_prof_n1_native.txt > gpurun_out/c3_bench_native.log 2>&1
echo "rc=$?" >> gpurun_out/c3_bench_native.log
CB200_GEMM_BACKEND=cublas timeout 600 python bench.py --gpus 1 --steps 4 --warmup 3 --no-e2e > gpurun_out/c3_bench_cublas.log 2>&1
echo "rc=$?" >> gpurun_out/c3_bench_cublas.log
tail -4 gpurun_out/c3_pytest_gpu.log
python - <<'PY'
import json
for f in ("gpurun_out/c3_gemm_m4096_split.jsonl","gpurun_out/c3_gemm_m4096_nosplit.jsonl"):
    print(f)
    for l in open(f):
        try: r=json.loads(l)
        except Exception: continue
        print(r["name"], r["shape"], {k:(round(r[k]["ours_tflops"]), round(r[k]["cublas_tflops"])) for k in ("nt","nn","tn")})
PY
tail -c 600 gpurun_out/c3_bench_native.log; echo; tail -c 600 gpurun_out/c3_bench_cublas.log
