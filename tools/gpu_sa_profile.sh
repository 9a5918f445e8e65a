#!/bin/bash
# usage: tools/gpu_sa_profile.sh <tag>  -- ncu --set full of the fused attend kernel (3rd launch), details + source pages
TAG=${1:-sa}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slot_attend_fused_kernel -s 2 -c 1 -f \
  -o gpurun_out/ncu_${TAG} python tools/sa_profile.py > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
ls -la gpurun_out/
