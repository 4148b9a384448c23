"""End-to-end parity of the other two BASELINE architectures through the drop-in
``VolSeg2dPredictor``: U-Net++ / ResNeXt-50_32x4d with 6 classes
(BASELINE cfg4) and DeepLabV3+ / ResNet-50 with 4 classes (cfg5), 3-way prediction of a
ragged volume, against the fp32 CPU oracle on the same seeded weights (BN statistics
randomised).  For DeepLabV3+ this is the test of the head kernel's fused bilinear x4
``align_corners=True`` up-sampling (kernels_simple.cu logit_at) against
``nn.UpsamplingBilinear2d`` (oracle/smp_models.py), of the global-average-pool branch, the
bilinear decoder up-sampling and the dilated depthwise convolutions at rates 12 / 24 / 36.

Criteria: winning probability within 2e-2; every label disagreement at a voxel whose
reference margin (winning label vs the best competing label over all directions) is below
2e-2; the agreement itself is printed (random-init weights: every voxel is a near-tie).
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import predict_oracle as po
from oracle.make_golden import structured_volume
from oracle.smp_models import make_random_model

pytestmark = pytest.mark.gpu

PROB_TOL = 2e-2
SETTINGS = dict(quality="medium", output_probs=False, clip_data=False, st_dev_factor=2.575,
                data_hdf5_path="/data", cuda_device=0, downsample=False, one_hot=False, prediction_axis="Z")

ARCHS = [
    ("U_NET_PLUS_PLUS", "unetplusplus", "resnext50_32x4d", 6, (20, 45, 70)),
    ("DEEPLABV3_PLUS", "deeplabv3plus", "resnet50", 4, (19, 70, 100)),
    ("U_NET", "unet", "resnet50", 2, (12, 33, 61)),
    # the first of the remaining smp model types (model_2d.py:21-38), plain atrous ASPP at
    # output stride 8, head bilinear x8 -- runs on the existing kernel classes
    ("DEEPLABV3", "deeplabv3", "resnet34", 3, (14, 70, 100)),
    ("DEEPLABV3", "deeplabv3", "resnet50", 2, (9, 45, 64)),
]


def _predictor(tmp_path, mt_name, oracle_model, encoder, classes):
    import volume_segmantics.utilities.base_data_utils as utils
    from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor

    struc = {"type": utils.ModelType[mt_name], "encoder_name": encoder, "encoder_weights": None,
             "in_channels": 1, "classes": classes}
    path = tmp_path / f"{mt_name}.pytorch"
    torch.save({"model_state_dict": oracle_model.state_dict(), "model_struc_dict": struc, "label_codes": {}}, path)
    return VolSeg2dPredictor(str(path), SimpleNamespace(**SETTINGS))


@pytest.mark.parametrize("mt_name,arch,encoder,classes,shape", ARCHS)
def test_three_way_end_to_end(tmp_path, mt_name, arch, encoder, classes, shape):
    oracle_model = make_random_model(arch, encoder, classes, seed=0)
    pred = _predictor(tmp_path, mt_name, oracle_model, encoder, classes)
    oracle = po.OraclePredictor(oracle_model, classes)
    vol = structured_volume(shape, 7 + classes)
    labels, probs = pred._predict_3_ways_max_probs(vol)
    want_l, want_p = oracle.predict_3_ways_max_probs(vol)
    assert labels.shape == shape and labels.dtype == np.uint8 and probs.dtype == np.float16
    perr = np.abs(probs.astype(np.float32) - want_p.astype(np.float32))
    agree = labels == want_l
    cb = np.sort(oracle.class_best_over_directions(vol, (0, 1, 2)), axis=0)
    margin = cb[-1] - cb[-2]
    worst = margin[~agree].max() if (~agree).any() else 0.0
    print(f"[e2e {mt_name}/{encoder} C={classes} {shape}] agreement {agree.mean():.5f} max prob err {perr.max():.5f} "
          f"largest reference margin at a disagreement {worst:.5f}")
    assert perr.max() < PROB_TOL
    assert worst < PROB_TOL
    assert agree.mean() >= 0.99


def test_deeplab_single_axis_full_probabilities(tmp_path):
    """All class probabilities (not just the winner) after the fused bilinear head, per axis."""
    from volume_segmantics.utilities.base_data_utils import Axis

    oracle_model = make_random_model("deeplabv3plus", "resnet50", 4, seed=1)
    pred = _predictor(tmp_path, "DEEPLABV3_PLUS", oracle_model, "resnet50", 4)
    oracle = po.OraclePredictor(oracle_model, 4)
    vol = structured_volume((9, 61, 93), 5)  # pads 3 and 3: crop shifted by one pixel against the pad
    for axis in (Axis.Z, Axis.X):
        labels, probs = pred._predict_single_axis(vol, axis=axis)
        want_l, want_p, full = oracle.predict_single_axis(vol, True, axis.value, return_full=True)
        perr = np.abs(probs.astype(np.float32) - np.ascontiguousarray(want_p).astype(np.float32)).max()
        top2 = np.sort(full, axis=1)[:, -2:]
        margin = po.rotate_array_to_axis(top2[:, 1] - top2[:, 0], axis.value)
        bad = labels != want_l
        print(f"[deeplab axis {axis.name}] agreement {1 - bad.mean():.5f} max prob err {perr:.5f}")
        assert perr < PROB_TOL
        assert not bad.any() or margin[bad].max() < PROB_TOL
