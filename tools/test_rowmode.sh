#!/bin/bash
# which descriptor convention does the hardware want for row-shifted SW128 views?
for bo in 1 0; do
  echo "=== FMDM_CONV_DESC_BASE_OFFSET=$bo"
  FMDM_CONV_DESC_BASE_OFFSET=$bo timeout 300 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -k "conv" --tb=line 2>&1 | tail -8
done
