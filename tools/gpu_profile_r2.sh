#!/bin/bash
# Round-2 profile set (one GPU):  gpurun --timeout 1500 -- 'bash tools/gpu_profile_r2.sh'
#  1. launch list + DRAM traffic of EVERY kernel of one SlotAttention + UNet evaluation at B=256 (time, dram bytes read/written)
#  2. the same for one full-model training step at B=64
#  3. --set full of the dominant kernels (first launches), raw CSV pages
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tc.sum"
timeout 500 ncu --profile-from-start off --metrics $M --clock-control none --csv \
  --log-file gpurun_out/r2c_traffic_inference_b256.csv python bench.py --profile-once --batch 256 > gpurun_out/r2c_prof1.log 2>&1 < /dev/null
tail -1 gpurun_out/r2c_prof1.log
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2c_traffic_train_full_b64.csv python bench.py --profile-train-once --train-batch 64 > gpurun_out/r2c_prof2.log 2>&1 < /dev/null
tail -1 gpurun_out/r2c_prof2.log
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"gemm_kernel|slot_attend_fused_kernel|attention_tc_kernel|groupnorm_apply_pack_fused" -c 40 -f -o /tmp/full_r2c \
  python bench.py --profile-once --batch 256 > gpurun_out/r2c_prof3.log 2>&1 < /dev/null
tail -1 gpurun_out/r2c_prof3.log
timeout 120 ncu -i /tmp/full_r2c.ncu-rep --page raw --csv > gpurun_out/r2c_ncu_full_top_kernels_b256_raw.csv 2>/dev/null < /dev/null
ls -la gpurun_out | grep r2c
