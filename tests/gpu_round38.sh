#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r01_gpu_tests_v3.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== slicer/head"; timeout 300 python tests/slicer_bench.py 2>&1 | tail -5 | tee gpurun_out/r01_slicer_head_bw.txt
echo "== layers"; timeout 600 python tests/layer_profile.py 1024 64 > gpurun_out/r01_layers_v20.txt 2>&1; tail -2 gpurun_out/r01_layers_v20.txt
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v15.json 2> gpurun_out/bench_v15.err; tail -3 gpurun_out/bench_v15.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v15.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_stage_ms_per_step'], d['cpu_baseline']['value'])
PY
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400
