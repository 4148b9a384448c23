#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== slicer/head"; timeout 300 python tests/slicer_bench.py 2>&1 | tail -6 | tee gpurun_out/r01_slicer_head_bw.txt
echo "== ncu slicer"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:slicer -s 4 -c 4 -f -o gpurun_out/r01_slicer python tests/slicer_bench.py > gpurun_out/ncu_slicer.log 2>&1; tail -1 gpurun_out/ncu_slicer.log
