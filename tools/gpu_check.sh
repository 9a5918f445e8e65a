#!/bin/bash
# usage: tools/gpu_check.sh  -- GEMM variant tests first (isolated process), then the full GPU suite, then a short bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "variants or scalar" 2>&1 | tail -15
echo "=== full suite"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15
echo "=== bench"
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -3
