#!/bin/bash
# One GPU round trip: base kernels, tcgen05 probes, whole network, end to end.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.log 2>&1
echo "== base tests" ; timeout 900 python -m pytest tests/test_slicer_gpu.py tests/test_merge_gpu.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/t_base.log
: > gpurun_out/probes.log
for grp in 0,1,2 3,4,5,6 7,8,9 10,11,12,13 14,15,16,17,18,19,20; do
  timeout 240 python tests/bringup_gpu.py --probe $grp >> gpurun_out/probes.log 2>&1 || echo "group $grp exit $?" >> gpurun_out/probes.log
done
echo "== probes"; grep -E "probe|exit|Error|error" gpurun_out/probes.log | tail -60
echo "== net"; timeout 900 python tests/bringup_gpu.py > gpurun_out/net.log 2>&1; tail -12 gpurun_out/net.log
echo "== net tests"; timeout 1200 python -m pytest tests/test_network_gpu.py tests/test_predictor_gpu.py -m gpu -q 2>&1 | tail -30 | tee gpurun_out/t_net.log
