#!/bin/bash
# TMA-store epilogue + multi-stage accumulators in conv_halo_kernel
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== layers (default)"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v10.txt | grep -v "layer3\|layer4"
echo "== layers (tma_epilogue=0)"; VSB_FLAGS=tma_epilogue=0 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep -v "layer3\|layer4" | tee gpurun_out/layers_v10_noepi.txt
echo "== layers (halo_a_stages=4)"; VSB_FLAGS=halo_a_stages=4 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep -v "layer3\|layer4" | tee gpurun_out/layers_v10_a4.txt | tail -25
