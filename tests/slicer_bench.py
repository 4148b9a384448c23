"""HBM throughput of the slicer and head kernels per direction type (run by hand under gpurun):
algorithmic bytes (3 B / padded pixel for the slicer, 32 B / pixel for the head at C = 4) over the
CUDA-event time of their launches."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nsl = 128
eng = Engine(0)
model = B200SegmentationModel("U_NET", "resnet34", 4)
eng.load_model(model)
vol = np.random.default_rng(0).integers(0, 256, (size, size, size), dtype=np.uint8)
eng.set_volume(vol)
for d, name in ((0, "Z (rows kernels)"), (1, "Y (rows kernels)"), (2, "X (x-plane kernels)"), (4, "rot90 Y"), (5, "rot90 X")):
    eng.predict_range(d, 0, nsl)
    eng.synchronize()
    eng.set_profiling(True)
    eng.predict_range(d, 0, nsl)
    eng.synchronize()
    st = eng.stage_times()
    eng.set_profiling(False)
    px = nsl * size * size
    s_ms, s_n = st["slicer"]
    h_ms, h_n = st["head"]
    print(f"direction {d:2d} {name:20s}: slicer {s_ms / s_n * 1e3:7.1f} us/launch = {3 * px / s_ms / 1e6:7.0f} GB/s ; "
          f"head {h_ms / h_n * 1e3:7.1f} us/launch = {32 * px / h_ms / 1e6:7.0f} GB/s")
