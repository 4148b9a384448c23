#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== layers (default)"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v13.txt | grep -v "layer3\|layer4"
echo "== layers (epi_groups=0)"; VSB_FLAGS=epi_groups=0 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "layer1.1\|blocks.2.conv2\|blocks.4\|segmentation_head\|total"
echo "== layers (tma_epilogue=0)"; VSB_FLAGS=tma_epilogue=0 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | grep "layer1.1\|blocks.2.conv2\|blocks.4\|segmentation_head\|total"
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; tail -3 gpurun_out/bench_v10.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v10.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['measured'], d['roofline']['other_stage_ms_per_step'])
PY
