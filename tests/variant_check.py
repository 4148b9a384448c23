"""Checks of ONE library build variant, run as a script with VSB200_VARIANT set (the variant is chosen
when the library is loaded, so it cannot change inside a pytest process):
    VSB200_VARIANT=bf16 python tests/variant_check.py
Slicer and injected merge bit-exact in the variant's 16-bit format, network + end-to-end within the
BASELINE tolerances (probability 2e-2, margin clause); prints the label agreement."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from conftest import act_bits, act_tag  # noqa: E402
from oracle import make_golden as mg  # noqa: E402
from oracle import predict_oracle as po  # noqa: E402
from oracle.smp_models import make_random_model  # noqa: E402
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel  # noqa: E402


def main():
    tag = act_tag()
    eng = Engine(0)
    # slicer, bit-exact
    vol = mg.synth_volume((9, 61, 33), 102)
    eng.set_volume(vol)
    for d in range(12):
        got = eng.slice_batch(d, 0, eng.geometry(d).S)
        assert np.array_equal(got, act_bits(po.slicer_oracle(vol, d))), f"slicer direction {d}"
    # injected merge, bit-exact
    z = np.load(ROOT / "tests" / "golden" / "merge_injected.npz")
    shape = z["high_labels"].shape
    eng.set_volume(np.zeros(shape, np.uint8))
    for d in range(12):
        eng.merge_injected(d, z[f"in_probs_{d}"], z[f"in_labels_{d}"])
    lab, prb = eng.fetch()
    assert np.array_equal(lab, z["high_labels"]) and np.array_equal(prb.view(np.uint16), z["high_probs"])
    # network + end to end
    oracle = make_random_model("unet", "resnet34", 4, seed=0)
    model = B200SegmentationModel("U_NET", "resnet34", 4)
    model.load_state_dict(oracle.state_dict())
    vol = mg.structured_volume((20, 40, 45), 3)
    eng.load_model(model)
    eng.set_volume(vol)
    eng.predict(0b111, True)
    labels, probs = eng.fetch()
    ora = po.OraclePredictor(oracle, 4)
    want_l, want_p = ora.predict_3_ways_max_probs(vol)
    perr = np.abs(probs.astype(np.float32) - want_p.astype(np.float32)).max()
    bad = labels != want_l
    cb = np.sort(ora.class_best_over_directions(vol, range(3)), axis=0)
    worst = (cb[-1] - cb[-2])[bad].max() if bad.any() else 0.0
    print(f"[variant {tag}] slicer + merge bit-exact; 3-way random-init: agreement {1 - bad.mean():.5f} "
          f"max prob err {perr:.5f} largest reference margin at a disagreement {worst:.5f}")
    assert perr < 2e-2 and worst < 2e-2
    eng.close()
    print(f"VARIANT {tag} OK")


if __name__ == "__main__":
    main()
