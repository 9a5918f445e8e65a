#!/bin/bash
# usage: tools/gpu_sa.sh -- fused Slot-Attention kernel: op parity, module parity vs golden, timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "slot_attend" 2>&1 | tail -25
echo "=== module tests"
timeout 300 python -m pytest tests/test_modules_gpu.py -q -m gpu -k "slot_attention" 2>&1 | tail -25
echo "=== sa bench"
timeout 300 python tools/sa_bench.py --batch 64 256 2>&1 | tail -8 | tee gpurun_out/sa_bench.log
