#!/bin/bash
# usage: tools/gpu_profile.sh <tag> [kernel-regex] [count]
#   launch list of one steady-state SA + UNet pass, then --set full of the selected kernels (raw CSV always comes back;
#   the .ncu-rep only if it fits gpurun's 64 MiB return limit).  Every step runs under its own timeout, stdin closed.
TAG=${1:-r1}
REGEX=${2:-gemm_kernel|slot_attend_fused_kernel|attention_tc_kernel}
COUNT=${3:-16}
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --profile-once > gpurun_out/prof_${TAG}.log 2>&1 < /dev/null
tail -2 gpurun_out/prof_${TAG}.log
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"${REGEX}" -c ${COUNT} -f -o /tmp/full_${TAG} \
  python bench.py --profile-once > gpurun_out/full_${TAG}.log 2>&1 < /dev/null
tail -2 gpurun_out/full_${TAG}.log
timeout 120 ncu -i /tmp/full_${TAG}.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_raw.csv 2>/dev/null < /dev/null
SZ=$(stat -c %s /tmp/full_${TAG}.ncu-rep 2>/dev/null || echo 999999999)
if [ "$SZ" -lt 45000000 ]; then cp /tmp/full_${TAG}.ncu-rep gpurun_out/; fi
ls -la gpurun_out | tail -6
