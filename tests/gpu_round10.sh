#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== archs"; timeout 900 python tests/arch_timing.py 2>&1 | tail -5 | tee gpurun_out/arch_timing.log
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; tail -3 gpurun_out/bench_v7.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['measured'], d['roofline']['other_stage_ms_per_step'])
PY
