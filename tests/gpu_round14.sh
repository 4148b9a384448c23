#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== layers (default)"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v11.txt | grep -v "layer3\|layer4"
echo "== layers (epi_groups=0)"; VSB_FLAGS=epi_groups=0 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "layer1.1\|layer2.1\|blocks.2.conv2\|blocks.4\|segmentation_head\|total"
for f in 7 3; do
echo "== layers (halo_dbg=$f)"; VSB_FLAGS=halo_dbg=$f timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "layer1.1\|layer2.1\|layer3.1\|blocks.2.conv2\|blocks.4\|segmentation_head"
done
