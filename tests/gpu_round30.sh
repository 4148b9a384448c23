#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== layers"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/r01_layers_v19.txt | grep "blocks\|total\|rror"
echo "== layers mma_warps=1"; VSB_FLAGS=mma_warps=1 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "blocks.2\|blocks.3\|total\|rror"
echo "== archs"; timeout 900 python tests/arch_timing.py 2>&1 | tail -4
