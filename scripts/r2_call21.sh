#!/bin/bash
# round-2 call 21 (1 GPU): evidence for the two newest kernels: flash forward timing (final) and ncu of the tensor-core
# decode kernel and of the final flash forward
mkdir -p gpurun_out; export PYTHONPATH=$PWD:$PYTHONPATH
timeout 80 python scripts/bench_flash_attn.py > gpurun_out/c21_flash_bench.log 2>&1
timeout 70 ncu --set full --clock-control none --import-source on -k regex:paged_decode_mma_kernel -s 2 -c 1 -f -o gpurun_out/ncu_paged_decode_mma python scripts/ncu_targets.py decode > gpurun_out/c21_ncu1.log 2>&1
ncu -i gpurun_out/ncu_paged_decode_mma.ncu-rep --page raw --csv > gpurun_out/ncu_paged_decode_mma_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_paged_decode_mma.ncu-rep --page details --csv > gpurun_out/ncu_paged_decode_mma_details.csv 2>/dev/null
timeout 70 ncu --set full --clock-control none --import-source on -k regex:flash_fwd_kernel -s 2 -c 1 -f -o gpurun_out/ncu_flash_fwd_v3 python scripts/ncu_targets.py flash_fwd > gpurun_out/c21_ncu2.log 2>&1
ncu -i gpurun_out/ncu_flash_fwd_v3.ncu-rep --page raw --csv > gpurun_out/ncu_flash_fwd_v3_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_flash_fwd_v3.ncu-rep --page details --csv > gpurun_out/ncu_flash_fwd_v3_details.csv 2>/dev/null
grep FLASH gpurun_out/c21_flash_bench.log | cut -c1-260; ls -la gpurun_out/ncu_paged_decode_mma* gpurun_out/ncu_flash_fwd_v3* | awk '{print $5, $9}'
