#!/bin/bash
# round-1 session 3: S2D tail + fast head/slicer kernels
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== layers"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/layers_v9.txt | tail -60
echo "== layers (no s2d)"; VSB200_S2D_TAIL=0 timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tail -8
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; tail -3 gpurun_out/bench_v9.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v9.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['measured'], d['roofline']['other_stage_ms_per_step'])
PY
echo "== ncu layer1"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 0 -c 2 -f -o gpurun_out/r01_l1conv python tests/layer_profile.py 1024 16 16 > gpurun_out/ncu_l1.log 2>&1; tail -2 gpurun_out/ncu_l1.log
echo "== ncu s2d tail"; timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 32 -c 2 -f -o gpurun_out/r01_s2dtail python tests/layer_profile.py 1024 16 16 > gpurun_out/ncu_s2d.log 2>&1; tail -2 gpurun_out/ncu_s2d.log
ls -la gpurun_out
