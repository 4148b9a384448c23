#!/bin/bash
mkdir -p gpurun_out
for v in f16 bf16; do
  echo "===== variant $v"
  VSB200_VARIANT=$v timeout 600 python tests/diag_e2e.py 2>&1 | tail -12
  VSB200_VARIANT=$v timeout 600 python tests/bringup_gpu.py 2>&1 | grep net
  VSB200_VARIANT=$v timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
done
