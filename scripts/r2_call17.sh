#!/bin/bash
# round-2 call 17 (2 GPUs): the driver's round-end checks on the final tree: full `pytest -m gpu`, then smoke()
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 420 python -m pytest tests/ -x -q -m gpu --timeout 240 -p no:cacheprovider > gpurun_out/c17_pytest_gpu.log 2>&1
echo "pytest_gpu rc=$?" >> gpurun_out/c17_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c17_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/c17_smoke.log
tail -8 gpurun_out/c17_pytest_gpu.log | cut -c1-300; tail -4 gpurun_out/c17_smoke.log | cut -c1-300
