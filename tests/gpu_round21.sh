#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== layers (default)"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v15.txt | grep "conv1 \|maxpool\|layer1.0\|blocks\|head\|total\|slicer"
