#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$" | tail -40 | tee gpurun_out/t_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench 512"; timeout 600 python bench.py --size 512 --steps 1 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench512.log
echo "== bench 1024"; timeout 1200 python bench.py 2>&1 | tail -3 | tee gpurun_out/bench1024.log
