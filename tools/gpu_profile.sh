#!/bin/bash
# usage: tools/gpu_profile.sh <tag> [kernel-regex] [count]
#   launch list of one steady-state SA + UNet pass, then --set full of the selected kernels (raw CSV always comes back;
#   the .ncu-rep only if it fits gpurun's 64 MiB return limit)
TAG=${1:-r1}
REGEX=${2:-gemm_kernel|slot_attend_kernel|groupnorm_stats|attention_pack}
COUNT=${3:-24}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_${TAG}.csv python bench.py --profile-once > gpurun_out/prof_${TAG}.log 2>&1
tail -2 gpurun_out/prof_${TAG}.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"${REGEX}" -c ${COUNT} -f -o /tmp/full_${TAG} \
  python bench.py --profile-once > gpurun_out/full_${TAG}.log 2>&1
tail -2 gpurun_out/full_${TAG}.log
ncu -i /tmp/full_${TAG}.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_raw.csv 2>/dev/null
SZ=$(stat -c %s /tmp/full_${TAG}.ncu-rep)
if [ "$SZ" -lt 45000000 ]; then cp /tmp/full_${TAG}.ncu-rep gpurun_out/; fi
ls -la gpurun_out
