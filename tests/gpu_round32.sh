#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== slicer/head"; timeout 300 python tests/slicer_bench.py 2>&1 | tail -6
echo "== bench"; timeout 1200 python bench.py --no-cpu --steps 1 --warmup 2 > gpurun_out/bench_v14.json 2> gpurun_out/bench_v14.err; tail -3 gpurun_out/bench_v14.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v14.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_stage_ms_per_step'])
PY
