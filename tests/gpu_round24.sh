#!/bin/bash
for i in 1 2 3; do
echo "== layers run $i"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "total\|Error\|error" | tail -3
done
for i in 1 2 3; do
echo "== layers sync_each run $i"; VSB_FLAGS=sync_each=1 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "total\|Error\|error" | tail -3
done
