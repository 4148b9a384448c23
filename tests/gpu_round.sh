#!/bin/bash
# Full validation on a B200 box (run as: gpurun --timeout 2400 -- 'bash tests/gpu_round.sh').
# Writes the artefacts that get copied into profiles/.
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/gpu_tests.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== slicer/head bandwidth"; timeout 300 python tests/slicer_bench.py 2>&1 | tail -5 | tee gpurun_out/slicer_head_bw.txt
echo "== per-layer table"; timeout 600 python tests/layer_profile.py 1024 64 > gpurun_out/layers.txt 2>&1; tail -2 gpurun_out/layers.txt
echo "== other architectures"; timeout 900 python tests/arch_timing.py 2>&1 | tail -4 | tee gpurun_out/arch_timing.txt
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_stage_ms_per_step'])
PY
# ncu (only after the plain runs above exited 0):
#   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
#       --log-file gpurun_out/ncu_batch32_dram.csv python tests/layer_profile.py 1024 32
#   ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 13 -c 1 -f -o gpurun_out/conv_halo_layer3 \
#       python tests/layer_profile.py 1024 16 16
