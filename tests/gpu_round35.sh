#!/bin/bash
echo "== slicer tests"; timeout 600 python -m pytest tests/test_slicer_gpu.py tests/test_ragged_gpu.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do echo "== slicer/head run $i"; timeout 300 python tests/slicer_bench.py 2>&1 | tail -5; done
