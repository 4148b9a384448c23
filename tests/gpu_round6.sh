#!/bin/bash
mkdir -p gpurun_out
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; tail -c 3000 gpurun_out/bench_v3.json
echo "== bench small for ncu"; python bench.py --size 256 --steps 1 --warmup 3 --no-cpu > gpurun_out/plain256.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_256.csv python bench.py --size 256 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
