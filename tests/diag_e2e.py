"""Numbers, not asserts: end-to-end agreement of the engine with the oracle goldens."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle.smp_models import make_random_model  # noqa: E402
from volume_segmantics_b200 import _lib  # noqa: E402
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel  # noqa: E402

g = np.load(ROOT / "tests/golden/e2e_unet_r34.npz")
oracle = make_random_model("unet", "resnet34", 4, seed=0)
model = B200SegmentationModel("U_NET", "resnet34", 4)
model.load_state_dict(oracle.state_dict())
eng = Engine(0)
eng.load_model(model)
vol = g["volume"]
eng.set_volume(vol)
print("variant", _lib.VARIANT, "act", _lib.act_dtype())
for name, mask, kl, kp in [("lowY", 2, "low_y_labels", "low_y_probs"), ("medium", 7, "medium_labels", "medium_probs"),
                           ("high", 4095, "high_labels", "high_probs")]:
    eng.reset()
    eng.predict(mask, True)
    lab, prb = eng.fetch()
    wl, wp = g[kl], g[kp].view(np.float16).astype(np.float32)
    pe = np.abs(prb.astype(np.float32) - wp)
    ag = lab == wl
    print(f"{name}: label agree {ag.mean():.5f}  prob maxerr all {pe.max():.5f} agree-only {pe[ag].max():.5f} "
          f"prob fp16 bits equal {np.mean(prb.view(np.uint16) == g[kp]):.4f} pmax range {wp.min():.3f}..{wp.max():.3f}")
for d in (0, 1, 2):
    full = g[f"full_probs_d{d}"]
    top2 = np.sort(full, axis=1)[:, -2:]
    margin = top2[:, 1] - top2[:, 0]
    print(f"d{d}: ref margin quantiles", np.quantile(margin, [0.01, 0.1, 0.5, 0.9, 0.99]).round(5))
