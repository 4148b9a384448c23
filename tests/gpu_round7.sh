#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== archs"; timeout 900 python tests/arch_timing.py 2>&1 | tail -6 | tee gpurun_out/arch_timing.log
