#!/bin/bash
for b in 32 64 96; do
echo "== bench batch $b"; timeout 900 python bench.py --batch $b --steps 1 --warmup 2 --no-cpu > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; tail -2 gpurun_out/bench_b$b.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b$b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['other_stage_ms_per_step'])
PY
done
