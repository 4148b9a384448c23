#!/bin/bash
# Round-2 validation on ONE B200 (gpurun --timeout 3000 -- 'bash tests/gpu_round2.sh [stage ...]').
# Stages: tests bench configs layers ncu ncustem ncuel ncudw traffic deeplab multi multicfg (the last two need gpurun --gpus N).
mkdir -p gpurun_out
STAGES="${@:-tests bench configs layers ncu}"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
for st in $STAGES; do
case $st in
tests)
  echo "== gpu tests"; timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02_gpu_tests.log 2>&1
  echo "pytest rc=$?"; grep -E "passed|failed|error" gpurun_out/r02_gpu_tests.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/r02_gpu_tests.log | head -20
  echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
  ;;
bench)
  echo "== bench cfg3"; timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err
  echo "rc=$?"; tail -3 gpurun_out/r02_bench_cfg3.err
  python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_cfg3.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks','result_sha256')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['other_stage_ms_per_step'])
    print({k:(round(v['frac'],3), round(v['ms_per_step'],1)) for k,v in d['roofline_classes'].items()})
except Exception as e: print('bench parse failed', e)
PY
  ;;
configs)
  for c in cfg2 cfg4 cfg5; do
    echo "== bench $c"; timeout 1500 python bench.py --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
    echo "rc=$?"; tail -3 gpurun_out/r02_bench_$c.err
    python - $c <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/r02_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['other_stage_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
  done
  ;;
layers)
  echo "== per-layer table"; timeout 600 python tests/layer_profile.py 1024 64 > gpurun_out/r02_layers.txt 2>&1; tail -2 gpurun_out/r02_layers.txt
  echo "== other architectures"; timeout 900 python tests/arch_timing.py > gpurun_out/r02_arch_timing.txt 2>&1; tail -6 gpurun_out/r02_arch_timing.txt
  echo "== per-layer, DeepLabV3+/R50 at 2048x2048 (cfg5 x-plane slices) and U-Net++/ResNeXt-50 at 512x512"
  timeout 600 python tests/layer_profile.py 2048 32 0 DEEPLABV3_PLUS resnet50 4 > gpurun_out/r02_layers_deeplab.txt 2>&1; tail -3 gpurun_out/r02_layers_deeplab.txt
  timeout 600 python tests/layer_profile.py 512 128 0 U_NET_PLUS_PLUS resnext50_32x4d 6 > gpurun_out/r02_layers_unetpp.txt 2>&1; tail -3 gpurun_out/r02_layers_unetpp.txt
  ;;
ncu)
  # only after the identical plain command exited 0
  echo "== ncu stem / b3.conv1 / conv_tc"
  timeout 300 python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_plain.log 2>&1 && \
  timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:stem_pool_kernel' -s 1 -c 1 -f -o gpurun_out/r02_stem \
      python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_stem.log 2>&1
  echo "ncu stem rc=$?"
  timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:conv_halo2_kernel<\(int\)0, \(int\)64, \(int\)2>' -s 5 -c 1 -f -o gpurun_out/r02_b3conv1 \
      python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_b3.log 2>&1
  echo "ncu b3 rc=$?"
  timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:conv_tc_kernel' -s 6 -c 2 -f -o gpurun_out/r02_convtc \
      python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_convtc.log 2>&1
  echo "ncu conv_tc rc=$?"
  ls -la gpurun_out/*.ncu-rep | tail -5
  ;;
ncustem)
  timeout 300 python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_plain.log 2>&1 && \
  timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:stem_pool' -s 1 -c 1 -f -o gpurun_out/r02_stem_v3 \
      python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_stem.log 2>&1
  echo "ncu stem rc=$?"
  ;;
ncuel)
  # entry-list kernel: decoder.blocks.3.conv1 (4th launch of the kernel in a batch: b0, b1, b2, b3 conv1 after layer2/3/4.0.conv1)
  timeout 300 python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_plain.log 2>&1 && \
  timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:conv_halo_el_kernel' -c 8 -f -o gpurun_out/r02_el \
      python tests/layer_profile.py 1024 16 16 > gpurun_out/r02_ncu_el.log 2>&1
  echo "ncu el rc=$?"
  ncu -i gpurun_out/r02_el.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/r02_el_raw.csv 2>/dev/null
  head -c 3000 gpurun_out/r02_el_raw.csv
  ;;
ncudw)
  # DeepLabV3+ CUDA-core kernels: dilated depthwise (ASPP), bilinear x4, generic head
  timeout 300 python tests/layer_profile.py 2048 4 4 DEEPLABV3_PLUS resnet50 4 > gpurun_out/r02_ncu_plain_dl.log 2>&1 && \
  timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:dwconv3x3_tiled_kernel|upsample_kernel|head_kernel' -c 8 -f -o gpurun_out/r02_dl_simt \
      python tests/layer_profile.py 2048 4 4 DEEPLABV3_PLUS resnet50 4 > gpurun_out/r02_ncu_dl.log 2>&1
  echo "ncu dw rc=$?"
  ncu -i gpurun_out/r02_dl_simt.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct > gpurun_out/r02_dl_simt_raw.csv 2>/dev/null
  head -c 6000 gpurun_out/r02_dl_simt_raw.csv
  ;;
traffic)
  # DRAM bytes per launch of one 128-slice batch (the benchmarked launch configuration), one ncu pass
  timeout 300 python tests/layer_profile.py 1024 128 128 > gpurun_out/r02_ncu_plain128.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled --csv \
      --log-file gpurun_out/r02_ncu_batch128_dram.csv python tests/layer_profile.py 1024 128 128 > gpurun_out/r02_ncu_traffic.log 2>&1
  echo "ncu traffic rc=$?"; wc -l gpurun_out/r02_ncu_batch128_dram.csv
  ;;
deeplab)
  timeout 600 python tests/layer_profile.py 2048 32 0 DEEPLABV3_PLUS resnet50 4 > gpurun_out/r02_layers_deeplab.txt 2>&1; tail -3 gpurun_out/r02_layers_deeplab.txt
  ;;
multi)
  # needs gpurun --gpus N (N >= 2)
  N=$(nvidia-smi -L | wc -l); echo "== multi-GPU on $N GPUs"
  timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r02_multi_gpu_tests_${N}gpu.log
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r02_bench_cfg3_${N}gpu.json 2> gpurun_out/r02_bench_cfg3_${N}gpu.err
  echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_cfg3_${N}gpu.err
  python - $N <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/r02_bench_cfg3_{sys.argv[1]}gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','result_sha256')}, d['e2e'])
except Exception as e: print('bench parse failed', e)
PY
  ;;
multicfg)
  # the other BASELINE configurations on N GPUs (gpurun --gpus N)
  N=$(nvidia-smi -L | wc -l)
  for c in cfg4 cfg5; do
    echo "== bench $c on $N GPUs"
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --config $c --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_bench_${c}_${N}gpu.json 2> gpurun_out/r02_bench_${c}_${N}gpu.err
    echo "rc=$?"; tail -3 gpurun_out/r02_bench_${c}_${N}gpu.err
    python - $c $N <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/r02_bench_{sys.argv[1]}_{sys.argv[2]}gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','result_sha256')}, d['e2e']['value'])
except Exception as e: print('bench parse failed', e)
PY
  done
  ;;
esac
done
