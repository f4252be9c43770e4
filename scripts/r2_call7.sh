#!/bin/bash
# round-2 call 7 (2 GPUs): fused ring attention (first run) + timing, multi-GPU pytest tier, MoE EP=2 bench, reference-arm rows,
# smoke() with peers, re-measure of the round-1 plugin anomalies with proper warm-up
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
CB200_RING_ATTN_TIMING=1 CB200_RING_LOCAL_TOKENS=8192 NGPU=2 timeout 600 python tests/test_shardformer/test_ring_attention.py > gpurun_out/c7_ring2.log 2>&1
echo "ring2 rc=$?" >> gpurun_out/c7_ring2.log
timeout 900 python -m pytest tests/test_parallel tests/test_moe tests/test_shardformer/test_ring_attention.py -m gpu -q > gpurun_out/c7_pytest_multi.log 2>&1
echo "pytest_multi rc=$?" >> gpurun_out/c7_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 scripts/bench_moe_ep.py > gpurun_out/c7_moe_ep2.log 2>&1
echo "moe_ep2 rc=$?" >> gpurun_out/c7_moe_ep2.log
CB200_REF_TRACE=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 \
  bench.py --impl reference --gpus 2 --steps 3 --warmup 3 --layers 8 > gpurun_out/c7_bench_ref.log 2>&1
echo "bench_ref rc=$?" >> gpurun_out/c7_bench_ref.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/c7_smoke.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641"
B="examples/language/llama/benchmark.py -c llama-100m -b 4 -l 1024 -s 10 -i 4"
for spec in "zero2:-p zero2" "gemini:-p gemini" "pp2:-p 3d --pp 2 --mbs 1" "pp2int:-p 3d --pp 2 --mbs 1 --pp_style interleaved --n_chunks 2" "pp2zbv:-p 3d --pp 2 --mbs 1 --pp_style zbv --n_chunks 2"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout -k 10 300 $TR $B $args > gpurun_out/c7_smoke_$name.log 2>&1
  echo "$name rc=$? $(grep -a 'throughput' gpurun_out/c7_smoke_$name.log | tail -1)" >> gpurun_out/c7_plugins.log
done
grep -E "RING_TIMING|RING_ATTN_GPU_OK|rc=|Error" gpurun_out/c7_ring2.log | cut -c1-600 | tail -6
tail -6 gpurun_out/c7_pytest_multi.log; grep -E "MOE_EP|rc=|Error" gpurun_out/c7_moe_ep2.log | cut -c1-900
tail -c 1800 gpurun_out/c7_bench_ref.log; tail -4 gpurun_out/c7_smoke.log | cut -c1-400; cat gpurun_out/c7_plugins.log
