#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/bringup_gpu.py --probe 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20 2>&1 | grep -E "probe|Error|error" | tail -30
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== layers"; timeout 300 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v2.log | tail -60
