#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 2 4 3 7; do
echo "== layers (halo_dbg=$f)"; VSB_FLAGS=halo_dbg=$f timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "layer1.1\|layer2.1\|layer3.1\|blocks.2.conv2\|blocks.4\|segmentation_head"
done
