#!/bin/bash
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== deeplab"; timeout 300 python tests/diag_deeplab.py 2>&1 | tail -12
echo "== archs"; timeout 900 python tests/arch_timing.py 2>&1 | tail -6 | tee gpurun_out/arch_timing_v2.log
