"""Per-layer timing table of the U-Net/ResNet-34 forward on slices of size^2
(run by hand under gpurun): ms, TFLOP/s and algorithmic GB/s per conv."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nslices = int(sys.argv[2]) if len(sys.argv) > 2 else 64
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
arch = sys.argv[4] if len(sys.argv) > 4 else "U_NET"
encoder = sys.argv[5] if len(sys.argv) > 5 else "resnet34"
classes = int(sys.argv[6]) if len(sys.argv) > 6 else 4
model = B200SegmentationModel(arch, encoder, classes)
eng = Engine(0)
eng.load_model(model)
if batch:
    eng.set_batch(batch)
import os  # noqa: E402
for kv in filter(None, os.environ.get("VSB_FLAGS", "").split(",")):  # e.g. VSB_FLAGS=tma_epilogue=0,halo_a_stages=4
    k, v = kv.split("=")
    eng.set_flag(k, int(v))
vol = np.random.default_rng(0).integers(0, 256, (nslices, size, size), dtype=np.uint8)
eng.set_volume(vol)
eng.predict_range(0, 0, nslices)
eng.synchronize()
eng.set_profiling(True)
eng.predict_range(0, 0, nslices)
eng.synchronize()
spec = model.spec
times = eng.op_times(len(spec.layers))
px = nslices * size * size
tot_ms = tot_fl = 0
print(f"size {size} slices {nslices} batch {batch or 'auto'}")
print(f"{'layer':38s} {'Cin':>5s} {'Cout':>5s} k s {'res':>5s} {'ms':>8s} {'TFLOP/s':>8s} {'GB/s':>7s} {'launch':>6s}")
for L, (ms, n) in zip(spec.layers, times):
    if L.kind not in ("conv", "maxpool", "gap", "upsample") or n == 0:
        continue
    ds = spec.tensors[L.out].ds_log2
    if L.kind in ("gap", "upsample"):
        t_in = spec.tensors[L.srcs[0][0]]
        by = t_in.channels * 2 * (px / 4 ** t_in.ds_log2 if t_in.ds_log2 >= 0 else nslices)
        by += spec.tensors[L.out].channels * 2 * (px / 4 ** ds if ds >= 0 else nslices)
        tot_ms += ms
        print(f"{L.kind:38s} {t_in.channels:5d} {spec.tensors[L.out].channels:5d} - - {'1/' + str(2 ** ds) if ds >= 0 else '1x1':5s} {ms:8.3f} {0.0:8.1f} {by / ms / 1e6:7.0f} {n:6d}")
        continue
    opx = px / 4 ** ds if ds >= 0 else nslices
    if L.kind == "conv":
        fl = 2.0 * L.k * L.k * (L.cin // L.groups) * L.cout * opx
        by = 0
        for t, up in L.srcs:
            tds = spec.tensors[t].ds_log2
            by += spec.tensors[t].channels * 2 * (px / 4 ** tds if tds >= 0 else nslices)
        by += L.cout * (4 if spec.tensors[L.out].dtype else 2) * opx
        if L.res >= 0:
            by += L.cout * 2 * opx
    else:
        fl, by = 0, spec.tensors[L.srcs[0][0]].channels * 2 * px / 4 ** (ds - 1) * 1.25
    tot_ms += ms
    tot_fl += fl
    res = f"1/{2 ** ds}" if ds >= 0 else "1x1"
    print(f"{(L.name or L.kind):38s} {L.cin:5d} {L.cout:5d} {L.k} {L.stride} {res:5s} {ms:8.3f} {fl / ms / 1e9:8.1f} {by / ms / 1e6:7.0f} {n:6d}")
print(f"total conv+pool ms {tot_ms:.2f}  ->  {tot_fl / tot_ms / 1e9:.1f} TFLOP/s ; per slice {tot_ms / nslices * 1e3:.1f} us")
print({k: (round(v[0], 2), v[1]) for k, v in eng.stage_times().items()})
