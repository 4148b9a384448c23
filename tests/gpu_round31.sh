#!/bin/bash
echo "== new tests"; timeout 900 python -m pytest tests/test_engine_paths_gpu.py -m gpu -q -x 2>&1 | tail -8
echo "== slicer/head"; timeout 300 python tests/slicer_bench.py 2>&1 | tail -6
