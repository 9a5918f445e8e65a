#!/bin/bash
timeout 300 python -m pytest tests/test_backward_ops_gpu.py -q -m gpu -x 2>&1 | tail -5
for t in test_slot_attention_gradients_match_oracle test_unet_small_gradients_match_oracle test_unet_full_gradients_match_reference_golden test_unet_dropout_train_mode test_denoise_loss_end_to_end; do
  echo "=== $t"
  timeout 400 python -m pytest tests/test_training_gpu.py -q -m gpu -x -s -k "$t" 2>&1 | grep -v "^$" | grep -E "passed|failed|Error|error|assert|worst|^E |line [0-9]+" | head -24
done
