#!/bin/bash
# usage: tools/gpu_attn.sh -- tensor-core attention: op parity, UNet module parity, per-launch timing tc vs CUDA-core
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -15
echo "=== unet module tests"
timeout 600 python -m pytest tests/test_modules_gpu.py -q -m gpu -k "unet or dpm" 2>&1 | tail -8
echo "=== attention timing"
timeout 300 python tools/attn_bench.py 2>&1 | tail -12 | tee gpurun_out/attn_bench.log
