#!/bin/bash
for f in 7 15 23 39 63 56 1 2; do
echo "== layers (halo_dbg=$f)"; VSB_FLAGS=halo_dbg=$f timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "layer1.1\|blocks.2.conv2\|blocks.4.conv2\|segmentation_head"
done
