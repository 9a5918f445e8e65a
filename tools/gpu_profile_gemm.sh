#!/bin/bash
# usage: tools/gpu_profile_gemm.sh [count] -- ncu --set full of the first <count> gemm_kernel launches of one UNet evaluation at B=256
COUNT=${1:-60}
mkdir -p gpurun_out
timeout 500 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_kernel -c ${COUNT} -f -o /tmp/full_gemm \
  python tools/profile_sampler.py > gpurun_out/full_gemm.log 2>&1 < /dev/null
tail -1 gpurun_out/full_gemm.log
timeout 120 ncu -i /tmp/full_gemm.ncu-rep --page raw --csv > gpurun_out/full_gemm_b256_raw.csv 2>/dev/null < /dev/null
ls -la gpurun_out/full_gemm_b256_raw.csv
