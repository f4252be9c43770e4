#!/bin/bash
# round-2 call 20 (1 GPU): bench.py contract run on the final tree (N = 1, short)
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 230 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/c20_bench_n1.log 2>&1
echo "bench rc=$?" >> gpurun_out/c20_bench_n1.log
tail -c 2500 gpurun_out/c20_bench_n1.log
