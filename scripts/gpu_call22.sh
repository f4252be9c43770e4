#!/bin/bash
# 1-GPU end-of-round sanity: gpu tests, smoke, bench N=1 (both arms), inference benchmark
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout -k 10 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_final.log | cut -c1-300
timeout -k 10 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_final.log | cut -c1-300
timeout -k 10 240 python bench.py --steps 3 --warmup 3 > gpurun_out/b1_final.log 2>&1; echo "bench1 rc=$?"; grep -a '"metric"' gpurun_out/b1_final.log | cut -c1-1800
timeout -k 10 150 python examples/inference/benchmark_llama.py -m llama3-8b --layers 8 -b 16 --in_len 512 --out_len 64 --cuda_graph > gpurun_out/infer_bench.log 2>&1; echo "infer rc=$?"; tail -6 gpurun_out/infer_bench.log | cut -c1-400
timeout -k 10 400 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/b1_ref.log 2>&1; echo "ref rc=$?"; grep -a '"impl"' gpurun_out/b1_ref.log | cut -c1-1500; grep -a "reference arm" gpurun_out/b1_ref.log | cut -c1-300
