#!/bin/bash
echo "== layers halo2_mma2=1"; VSB_FLAGS=halo2_mma2=1 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "blocks\|total\|rror"
