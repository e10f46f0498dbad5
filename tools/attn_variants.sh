#!/bin/bash
# Timing experiments on the attention backward's dQ MMAs (CARA_ATTN_VAR variants compute wrong numbers on purpose).
for v in 0 1 2 3 4; do
  CARA_NVCC_EXTRA="-DCARA_ATTN_DEBUG -DCARA_ATTN_VAR=$v" python -m cara_b200.build --force > /dev/null 2>&1
  echo "== variant $v"
  CARA_NVCC_EXTRA="-DCARA_ATTN_DEBUG -DCARA_ATTN_VAR=$v" python tools/attn_check.py 2>&1 | tail -2
done
