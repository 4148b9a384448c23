#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r01_gpu_tests_v2.log
echo "== layers"; timeout 600 python tests/layer_profile.py 1024 64 2>&1 | tee gpurun_out/r01_layers_v17.txt | grep "conv1 \|maxpool\|layer1.0\|blocks.3\|blocks.4\|head\|total\|slicer\|rror"
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; tail -3 gpurun_out/bench_v12.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v12.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['measured'], d['roofline']['other_stage_ms_per_step'], d['cpu_baseline'])
PY
echo "== ncu dram bytes per launch (one batch of 32 slices of 1024^2)"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_ncu_batch32_dram.csv python tests/layer_profile.py 1024 32 > gpurun_out/ncu_batch.log 2>&1; tail -2 gpurun_out/ncu_batch.log
echo "== ncu launch list of bench.py (512^3, 1 step)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_ncu_launches_512.csv python bench.py --size 512 --steps 1 --warmup 3 --no-cpu --no-profile > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
echo "== ncu full: layer3 conv_halo + S2D tail"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 13 -c 1 -f -o gpurun_out/r01_conv_halo_layer3_v2 python tests/layer_profile.py 1024 16 16 > gpurun_out/ncu_l3.log 2>&1; tail -1 gpurun_out/ncu_l3.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 32 -c 3 -f -o gpurun_out/r01_s2dtail_v2 python tests/layer_profile.py 1024 16 16 > gpurun_out/ncu_s2d.log 2>&1; tail -1 gpurun_out/ncu_s2d.log
ls -la gpurun_out | tail -12
