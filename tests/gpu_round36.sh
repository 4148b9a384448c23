#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:slicer_rows -s 2 -c 2 -f -o gpurun_out/r01_slicer_rows python tests/slicer_bench.py > gpurun_out/ncu_slicer.log 2>&1; tail -1 gpurun_out/ncu_slicer.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:slicer_xplane -s 1 -c 2 -f -o gpurun_out/r01_slicer_xplane python tests/slicer_bench.py > gpurun_out/ncu_slicer2.log 2>&1; tail -1 gpurun_out/ncu_slicer2.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:head_s2d -s 1 -c 3 -f -o gpurun_out/r01_head_s2d python tests/slicer_bench.py > gpurun_out/ncu_head.log 2>&1; tail -1 gpurun_out/ncu_head.log
