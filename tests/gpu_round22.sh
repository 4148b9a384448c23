#!/bin/bash
echo "== prof"; VSB_FLAGS=halo_prof=1 timeout 600 python tests/layer_profile.py 1024 32 2>&1 | grep "halo_prof" | awk '$3==3 || $3==4 || $3==43 || $3==44 || $3==45 || $3==46 || $3==47' | tail -7
