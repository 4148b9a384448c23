#!/bin/bash
mkdir -p gpurun_out
for b in 8 16 32; do echo "== batch $b"; timeout 300 python tests/layer_profile.py 1024 64 $b 2>&1 | grep -E "total|slicer"; done
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_stage_ms_per_step'])
PY
