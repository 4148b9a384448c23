"""Throughput sanity of the other BASELINE configs' architectures (run by hand under gpurun)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel, conv_macs_per_pixel  # noqa: E402

CASES = [("U_NET", "resnet34", 4, (512, 512, 512), 0b111, "cfg2 U-Net/R34 medium 512^3"),
         ("U_NET_PLUS_PLUS", "resnext50_32x4d", 6, (256, 256, 256), 0b111, "cfg4-like U-Net++/ResNeXt-50 medium 256^3"),
         ("DEEPLABV3_PLUS", "resnet50", 4, (128, 512, 512), 0b111, "cfg5-like DeepLabV3+/R50 medium 128x512x512"),
         ("U_NET", "resnet50", 2, (256, 256, 256), 0b111, "U-Net/R50 medium 256^3")]
only = sys.argv[1:] or None
eng = Engine(0)
for mt, enc, c, shape, mask, name in CASES:
    if only and mt not in only:
        continue
    model = B200SegmentationModel(mt, enc, c)
    eng.load_model(model)
    vol = np.random.default_rng(0).integers(0, 256, shape, dtype=np.uint8)
    eng.set_volume(vol)
    eng.predict(mask, True)
    eng.synchronize()
    eng.reset()
    eng.set_profiling(True)
    t0 = time.perf_counter()
    eng.predict(mask, True)
    eng.synchronize()
    dt = time.perf_counter() - t0
    st = eng.stage_times()
    eng.set_profiling(False)
    nvox = np.prod(shape)
    dirs = bin(mask).count("1")
    flops = 2 * conv_macs_per_pixel(model.spec) * nvox * dirs
    print(f"{name}: {dt * 1e3:.0f} ms, {nvox / dt / 1e6:.1f} Mvox/s, {flops / dt / 1e12:.0f} TFLOP/s overall; "
          + " ".join(f"{k}={v[0]:.0f}ms/{v[1]}" for k, v in st.items() if v[1]))
