#!/bin/bash
# First GPU call of the next round: run what was written blind at the end of round 1 (marker gpu_next), then time it.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
echo "=== gpu_next tests (one-launch slot update)"
timeout 300 python -m pytest tests -q -m gpu_next 2>&1 | tail -20 | tee gpurun_out/gpu_next_tests.log
echo "=== Slot Attention module timing: GEMM tail vs one-launch tail"
timeout 300 python tools/sa_bench.py --batch 4 16 64 256 2>&1 | tail -8 | tee gpurun_out/sa_bench_tail.log
echo "=== eager PyTorch bar on the same GPU (oracle arithmetic on CUDA; TF32 off / on)"
timeout 600 python tools/eager_gpu_bar.py 2>&1 | tail -6 | tee gpurun_out/eager_gpu_bar.log
echo "=== probe: fp16 main term + e4m3 correction terms vs three fp16 passes (DESIGN 8.0)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I slotdiffusion_b200/csrc tools/probes/umma_f8_probe.cu -o tools/probes/umma_f8_probe.bin \
  && timeout 120 ./tools/probes/umma_f8_probe.bin 2>&1 | tee gpurun_out/umma_f8_probe.log
echo "=== fused attend timeline at B=256 (which stage is long under full-chip load?)"
SA_B=256 timeout 120 python tools/sa_timeline.py 2>&1 | tail -40 | tee gpurun_out/sa_timeline_b256.log
echo "=== fused attend: experimental build (cheaper mbarrier polls, tail selects only in the partial group) -- parity, then timing"
# variants are prebuilt here (SDB_SF_EXPERIMENTAL=1 / SDB_GEMM_TIMING=1 python -m slotdiffusion_b200.build) and travel with the snapshot
EXP=$PWD/slotdiffusion_b200/libsdb200_sfexp.so
SDB_SF_EXPERIMENTAL=1 python -m slotdiffusion_b200.build > /dev/null 2>&1      # no-op when the prebuilt variant is up to date
SDB_LIB=$EXP timeout 300 python -m pytest tests/test_modules_gpu.py tests/test_ops_gpu.py -q -m gpu -k "slot_att" 2>&1 | tail -4 | tee gpurun_out/sf_experimental_tests.log
SDB_LIB=$EXP timeout 300 python tools/sa_bench.py --batch 64 256 2>&1 | tail -2 | tee gpurun_out/sa_bench_experimental.log
echo "=== GEMM issuer wait split (DESIGN 8.3): operands vs accumulator vs issuing"
TIM=$PWD/slotdiffusion_b200/libsdb200_gtiming.so
SDB_GEMM_TIMING=1 python -m slotdiffusion_b200.build > /dev/null 2>&1
SDB_LIB=$TIM timeout 300 python tools/gemm_wait_split.py 2>&1 | tail -10 | tee gpurun_out/gemm_wait_split.log
echo "=== full-model training step (reference's eager encoder / VQ-VAE encoder around the B200 modules)"
timeout 300 python tools/full_model_train_bench.py --batch 64 2>&1 | tail -2 | tee gpurun_out/full_model_train.log
timeout 300 python tools/full_model_train_bench.py --frames 3 --slots 15 --iters 2 --batch 16 2>&1 | tail -2 | tee gpurun_out/full_model_train_video.log
# second call (separate, ~6 GPU-min): source-level ncu reports to read offline with tools/ncu_wait_share.py / ncu_stalls.py
#   gpurun --timeout 900 -- 'bash tools/gpu_profile.sh r2a "gemm_kernel" 6; SA_B=256 bash tools/gpu_sa_profile.sh sa_b256'
#   sanitizer pass over the new kernel (memcheck + racecheck on the smallest cases; slow, keep it to a few tests):
#   gpurun --timeout 900 -- 'compute-sanitizer --tool memcheck python -m pytest tests/test_slot_update_gpu.py -q -m gpu_next -k "op_matches or ragged" 2>&1 | tail -15; compute-sanitizer --tool racecheck python -m pytest tests/test_slot_update_gpu.py -q -m gpu_next -k "op_matches" 2>&1 | tail -15'
