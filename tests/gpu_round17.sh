#!/bin/bash
for f in 0 63; do
echo "== prof (halo_dbg=$f)"; VSB_FLAGS=halo_dbg=$f,halo_prof=1 timeout 600 python tests/layer_profile.py 1024 32 2>&1 | grep "halo_prof" | awk '$3==3 || $3==4 || $3==12 || $3==20 || $3==43 || $3==46 || $3==47' | tail -7
done
