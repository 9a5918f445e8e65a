#!/bin/bash
# usage: tools/gpu_final.sh <tag> -- round-end validation: full GPU tests, smoke, default bench, B=256 sampler launch list
TAG=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu < /dev/null 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" < /dev/null 2>&1 | tail -2
timeout 400 python bench.py < /dev/null 2>gpurun_out/bench_${TAG}.err | tail -1 > gpurun_out/bench_${TAG}.json
python tools/show_bench.py gpurun_out/bench_${TAG}.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -c 430 \
  --log-file gpurun_out/launches_sampler_b256_${TAG}.csv python tools/profile_sampler.py > gpurun_out/prof_sampler_${TAG}.log 2>&1 < /dev/null
python tools/launch_summary.py gpurun_out/launches_sampler_b256_${TAG}.csv | head -24
