#!/bin/bash
mkdir -p gpurun_out
echo "== slicer tests"; timeout 600 python -m pytest tests/test_slicer_gpu.py tests/test_ragged_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "== slicer/head"; timeout 300 python tests/slicer_bench.py 2>&1 | tail -6
echo "== ncu slicer"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:slicer -c 6 -f -o gpurun_out/r01_slicer python tests/slicer_bench.py > gpurun_out/ncu_slicer.log 2>&1; tail -1 gpurun_out/ncu_slicer.log
