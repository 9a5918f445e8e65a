#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"groupnorm_apply_pack_fused_kernel|pack_nhwc_kernel|layernorm_pack_kernel" -c 14 -f -o /tmp/full_mem \
  python bench.py --profile-once --batch 64 > gpurun_out/full_mem.log 2>&1 < /dev/null
tail -1 gpurun_out/full_mem.log
timeout 120 ncu -i /tmp/full_mem.ncu-rep --page raw --csv > gpurun_out/full_mem_raw.csv 2>/dev/null < /dev/null
ls -la gpurun_out/full_mem_raw.csv
