#!/bin/bash
# compute-sanitizer targets for the single-GPU kernel tests (memcheck / racecheck / synccheck / initcheck).
# The reference has no sanitizer hooks; the flag protocols and smem pipelines here warrant them.
#   bash scripts/sanitize.sh memcheck  tests/test_kernels/test_elementwise_norm.py
#   bash scripts/sanitize.sh racecheck tests/test_kernels/test_gemm_tcgen05.py -k "128 and nt"
set -e
tool=${1:-memcheck}; shift || true
target=${@:-tests/test_kernels/test_elementwise_norm.py}
export PYTHONPATH=$PWD
mkdir -p gpurun_out
compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 99 --print-limit 20 \
  python -m pytest $target -x -q -m gpu 2>&1 | tee "gpurun_out/sanitize_${tool}.log" | tail -25
