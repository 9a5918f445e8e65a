#!/bin/bash
for t in test_conv3_backward test_conv3_stride2_backward test_groupnorm_backward test_layernorm_backward test_attention_backward test_pointwise_backward test_gru_backward test_slot_attend_backward; do
  echo "=== $t"
  timeout 300 python -m pytest tests/test_backward_ops_gpu.py -q -m gpu -x -s -k "$t" 2>&1 | grep -v "^$" | grep -E "passed|failed|Error|error|assert|timeout|sdb200|^E " | head -12
done
