"""The CPU oracle against the committed golden vectors (tests/golden, produced by
oracle/make_golden.py) and against the reference's own known answers."""
import hashlib
import json

import numpy as np
import torch

import torch as _t

from conftest import act_bits
from oracle import make_golden as mg
from oracle import predict_oracle as po


def test_slicer_oracle_digests(golden_dir):
    digests = json.loads((golden_dir / "slicer_digests.json").read_text())
    for si, shape in enumerate(mg.SLICER_SHAPES[:3]):
        vol = mg.synth_volume(shape, 100 + si)
        for d in range(12):
            ref = po.slicer_oracle(vol, d)
            rec = digests[f"{shape}|{d}"]
            assert list(ref.shape) == rec["shape"]
            assert hashlib.sha256(act_bits(ref, _t.bfloat16).tobytes()).hexdigest() == rec["sha256_bf16"]
            assert hashlib.sha256(act_bits(ref, _t.float16).tobytes()).hexdigest() == rec["sha256_f16"]


def test_slicer_oracle_is_reflect101_centre_pad():
    img = np.arange(5 * 7, dtype=np.uint8).reshape(5, 7)
    out = po.pad_slice(img)
    assert out.shape == (32, 32)
    top, left = int(27 / 2.0), int(25 / 2.0)
    assert np.array_equal(out[top:top + 5, left:left + 7], img)
    # repeated reflection without repeating the edge pixel
    assert out[top - 1, left] == img[1, 0] and out[top - 4, left] == img[4, 0] and out[top - 5, left] == img[3, 0]


def test_merge_oracle_golden_and_nested_equals_flat(golden_dir):
    z = np.load(golden_dir / "merge_injected.npz")
    shape = z["high_labels"].shape
    probs = {d: z[f"in_probs_{d}"] for d in range(12)}
    labels = {d: z[f"in_labels_{d}"] for d in range(12)}
    lab, prb = po.merge_injected_oracle(shape, list(range(12)), probs, labels)
    assert np.array_equal(lab, z["high_labels"]) and np.array_equal(prb.view(np.uint16), z["high_probs"])
    # the reference's nesting (3-way folds inside the 12-way fold) == flat first-max
    lc, pc = np.zeros((2, *shape), np.uint8), np.zeros((2, *shape), np.float16)
    for k in range(4):
        l3, p3 = po.merge_injected_oracle(shape, [3 * k, 3 * k + 1, 3 * k + 2], probs, labels)
        if k == 0:
            lc[0], pc[0] = l3, p3
        else:
            lc[1], pc[1] = l3, p3
            po.merge_vols_in_mem(pc, lc)
    assert np.array_equal(lc[0], lab) and np.array_equal(pc[0], prb)
    # skipping the four duplicate directions cannot change the result only when
    # their inputs really are duplicates; with independent injected data it may
    assert z["high_nodup_labels"].shape == shape


def test_network_oracle_golden(golden_dir):
    from oracle.smp_models import make_random_model

    z = np.load(golden_dir / "network_logits.npz")
    model = make_random_model("unet", "resnet34", 4, seed=0)
    vol = mg.structured_volume((2, 40, 70), 11)
    imgs = np.stack([po.preprocess_slice(vol[i]) for i in range(2)]).astype(np.float32)
    with torch.no_grad():
        logits = model(torch.from_numpy(imgs)[:, None]).numpy()
    assert np.allclose(logits, z["unet|resnet34|4"], atol=2e-4)


def test_oracle_predictor_contract():
    """dtype/shape contract of reference tests/test_vol_seg_2d_predictor.py:22-81."""
    from oracle.smp_models import make_random_model

    model = make_random_model("unet", "resnet18", 3, seed=1)
    pred = po.OraclePredictor(model, 3)
    vol = np.random.default_rng(0).integers(0, 256, (6, 33, 20), dtype=np.uint8)
    lab, prb = pred.predict_single_axis(vol, True, po.AXIS_X)
    assert lab.dtype == np.uint8 and prb.dtype == np.float16 and lab.shape == vol.shape == prb.shape
    lab, prb = pred.predict_single_axis(vol, False)
    assert prb is None
    oh = pred.predict_single_axis_to_one_hot(vol)
    assert oh.dtype == np.uint8 and oh.ndim == 4 and oh.shape[0] == 3
