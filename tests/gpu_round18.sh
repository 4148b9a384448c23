#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== layers (default)"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v12.txt | grep -v "layer3\|layer4"
echo "== layers (mma_warps=1)"; VSB_FLAGS=mma_warps=1 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "layer1.1\|layer2.1\|blocks.2.conv2\|blocks.4\|segmentation_head\|total"
echo "== prof"; VSB_FLAGS=halo_prof=1 timeout 600 python tests/layer_profile.py 1024 32 2>&1 | grep "halo_prof" | awk '$3==3 || $3==4 || $3==12 || $3==46 || $3==47' | tail -5
