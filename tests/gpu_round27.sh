#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== layers"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/r01_layers_v18.txt | grep "conv1 \|maxpool\|layer1\|layer2.1\|layer3.1\|layer4.1\|blocks\|head\|total\|slicer\|rror"
echo "== archs"; timeout 900 python tests/arch_timing.py 2>&1 | tail -6 | tee gpurun_out/arch_timing_v2.log
