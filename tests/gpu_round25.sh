#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for i in 1 2; do
echo "== layers run $i"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v16.txt | grep "conv1 \|maxpool\|layer1.0\|layer2.1\|blocks\|head\|total\|slicer\|rror"
done
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; tail -3 gpurun_out/bench_v11.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v11.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['measured'], d['roofline']['other_stage_ms_per_step'])
PY
