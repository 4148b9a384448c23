#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== layers (default)"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v16.txt | grep "conv1 \|maxpool\|layer1.0\|layer2.1\|blocks\|head\|total\|slicer"
echo "== prof"; VSB_FLAGS=halo_prof=1 timeout 600 python tests/layer_profile.py 1024 32 2>&1 | grep "halo_prof" | awk '$3==44 || $3==45 || $3==46 || $3==47' | tail -4
