"""Per-direction wall time of the DeepLabV3+ case (run by hand under gpurun)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel  # noqa: E402

eng = Engine(0)
model = B200SegmentationModel("DEEPLABV3_PLUS", "resnet50", 4)
eng.load_model(model)
vol = np.random.default_rng(0).integers(0, 256, (128, 512, 512), dtype=np.uint8)
eng.set_volume(vol)
for rep in range(3):
    for d in (0, 1, 2):
        eng.set_profiling(True)
        t0 = time.perf_counter()
        eng.predict(1 << d, True)
        t1 = time.perf_counter()
        eng.synchronize()
        t2 = time.perf_counter()
        st = eng.stage_times()
        eng.set_profiling(False)
        print(f"rep {rep} dir {d}: enqueue {1e3 * (t1 - t0):.1f} ms, total {1e3 * (t2 - t0):.1f} ms; kernels "
              + " ".join(f"{k}={v[0]:.1f}/{v[1]}" for k, v in st.items() if v[1]), flush=True)
spec = model.spec
times = eng.op_times(len(spec.layers))
for L, (ms, n) in zip(spec.layers, times):
    if n and ms > 5:
        print(f"  {L.kind:8s} {L.name:40s} {ms:9.2f} ms {n}")
