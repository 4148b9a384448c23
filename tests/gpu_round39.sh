#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== layers"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/r01_layers_v21.txt | grep "blocks\|total\|rror"
echo "== layers halo2_tma=0"; VSB_FLAGS=halo2_tma=0 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "blocks.*conv1\|blocks.3\|total\|rror"
echo "== archs"; timeout 900 python tests/arch_timing.py 2>&1 | tail -4
