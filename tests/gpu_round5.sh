#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for mb in 0 40 80 160; do echo "== layers sub_batch_mb=$mb"; VSB_SUB_BATCH_MB=$mb timeout 300 python tests/layer_profile.py 1024 64 2>&1 > gpurun_out/layers_sb$mb.log; grep -E "conv1  |maxpool|layer1.0|layer2.1.conv1|blocks.2|blocks.3|blocks.4|segmentation|total" gpurun_out/layers_sb$mb.log; done
