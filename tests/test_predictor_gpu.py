"""End to end through the reference-facing API (VolSeg2dPredictor /
VolSeg2DPredictionManager) against the CPU oracle.  BASELINE.json tolerances:
  * per-voxel (max) class probability within 2e-2 absolute,
  * label agreement >= 99.9 %, every disagreement at a voxel whose reference
    top-2 margin is below that tolerance.
Two weight sets are used:
  * "trained": the oracle after ~40 Adam steps on synthetic labels
    (oracle/train_synth.py) -- decisive like a real checkpoint; compared LIVE
    against the oracle run on the same in-memory weights; the 99.9 % bar applies.
  * "random-init" golden vectors (tests/golden/e2e_unet_r34.npz): softmax within
    0.004 of uniform, i.e. every voxel is on a decision boundary (reference margin
    < 2e-2 everywhere).  Probability tolerance and the margin clause apply in full.
    The agreement measures the rounding noise of the 16-bit activation chain (logit
    error about 1e-4 on logits of magnitude 0.07): measured 99.87 % single axis,
    99.83 % 3-way, 99.79 % 12-way (profiles/r01_gpu_tests_v4.log) -- BELOW the 99.9 %
    north_star asks for, and stated as such in DESIGN.md; RANDOM_INIT_FLOOR only keeps
    it from regressing.
For 3-way and 12-way results the margin clause uses the reference's margin between the
winning label and the best competing label over all merged directions
(OraclePredictor.class_best_over_directions).
Plus the dtype / shape contract of the reference's own GPU tests
(tests/test_vol_seg_2d_predictor.py:14-81, test_vol_seg_prediction_manager.py)."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import predict_oracle as po
from oracle.make_golden import structured_volume

pytestmark = pytest.mark.gpu

PROB_TOL = 2e-2
TRAINED_FLOOR = 0.999      # north_star
RANDOM_INIT_FLOOR = 0.997  # every voxel a near-tie: see the module docstring
SETTINGS = dict(quality="medium", output_probs=False, clip_data=False, st_dev_factor=2.575,
                data_hdf5_path="/data", cuda_device=0, downsample=False, one_hot=False, prediction_axis="Z")


def _save(model, path):
    import volume_segmantics.utilities.base_data_utils as utils

    struc = {"type": utils.ModelType.U_NET, "encoder_name": "resnet34", "encoder_weights": None,
             "in_channels": 1, "classes": 4}
    torch.save({"model_state_dict": model.state_dict(), "model_struc_dict": struc, "label_codes": {}}, path)
    return path


@pytest.fixture(scope="module")
def model_path(tmp_path_factory, unet_r34):
    return _save(unet_r34[0], tmp_path_factory.mktemp("model") / "test_model.pytorch")


@pytest.fixture(scope="module")
def predictor(model_path):
    from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor

    return VolSeg2dPredictor(str(model_path), SimpleNamespace(**SETTINGS))


@pytest.fixture(scope="module")
def trained(tmp_path_factory, trained_unet_r34):
    from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor

    path = _save(trained_unet_r34[0], tmp_path_factory.mktemp("model_t") / "trained.pytorch")
    vol = structured_volume((24, 40, 45), 77)
    return VolSeg2dPredictor(str(path), SimpleNamespace(**SETTINGS)), po.OraclePredictor(trained_unet_r34[0], 4), vol


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(golden_dir / "e2e_unet_r34.npz")


def _check(name, labels, probs, want_l, want_p, min_agree, ref_margin=None):
    assert labels.dtype == np.uint8 and probs.dtype == np.float16 and labels.shape == want_l.shape
    want_p = np.asarray(want_p).view(np.float16).astype(np.float32) if want_p.dtype == np.uint16 else want_p.astype(np.float32)
    perr = np.abs(probs.astype(np.float32) - want_p)
    agree = labels == want_l
    print(f"[{name}] label agreement {agree.mean():.5f}  max prob error {perr.max():.5f}  "
          f"ref p_max median {np.median(want_p):.3f}")
    # the winning probability is within tolerance everywhere (also where the labels differ:
    # then the two candidates were within tolerance of each other = the margin clause for merges)
    assert perr.max() < PROB_TOL, f"{name}: max prob error {perr.max()}"
    assert agree.mean() >= min_agree, f"{name}: label agreement {agree.mean():.5f} < {min_agree}"
    if ref_margin is not None and (~agree).any():
        assert (ref_margin[~agree] < PROB_TOL).all(), f"{name}: disagreement at a voxel with margin >= {PROB_TOL}"


def test_init_attributes(predictor):
    assert isinstance(predictor.model_file_path, Path) and isinstance(predictor.model, torch.nn.Module)
    assert predictor.model_device_num == 0 and predictor.num_labels == 4 and isinstance(predictor.label_codes, dict)


# ---------------------------------------------------------------- trained weights, 99.9 % bar
@pytest.mark.parametrize("axis_name", ["Z", "Y", "X"])
def test_trained_single_axis(trained, axis_name):
    from volume_segmantics.utilities.base_data_utils import Axis

    pred, oracle, vol = trained
    axis = Axis[axis_name]
    labels, probs = pred._predict_single_axis(vol, axis=axis)
    want_l, want_p, full = oracle.predict_single_axis(vol, True, axis.value, return_full=True)
    top2 = np.sort(full, axis=1)[:, -2:]
    margin = po.rotate_array_to_axis(top2[:, 1] - top2[:, 0], axis.value)
    _check(f"trained {axis_name}", labels, probs, np.ascontiguousarray(want_l), np.ascontiguousarray(want_p),
           TRAINED_FLOOR, margin)
    labels2, none = pred._predict_single_axis(vol, output_probs=False, axis=axis)
    assert none is None and np.array_equal(labels2, labels)


def _merged_margin(oracle, vol, dirs):
    cb = np.sort(oracle.class_best_over_directions(vol, dirs), axis=0)
    return cb[-1] - cb[-2]


def test_trained_three_ways(trained):
    pred, oracle, vol = trained
    labels, probs = pred._predict_3_ways_max_probs(vol)
    want_l, want_p = oracle.predict_3_ways_max_probs(vol)
    _check("trained 3-way", labels, probs, want_l, want_p, TRAINED_FLOOR, _merged_margin(oracle, vol, range(3)))


def test_trained_twelve_ways_and_one_hot(trained):
    pred, oracle, vol = trained
    labels, probs = pred._predict_12_ways_max_probs(vol)
    want_l, want_p = oracle.predict_12_ways_max_probs(vol)
    _check("trained 12-way", labels, probs, want_l, want_p, TRAINED_FLOOR, _merged_margin(oracle, vol, range(12)))
    votes = pred._predict_12_ways_one_hot(vol)
    want = oracle.predict_12_ways_one_hot(vol)
    assert votes.dtype == np.uint8 and votes.ndim == 4 and votes.shape == want.shape
    assert (votes.sum(0) == 12).all()
    moved = np.abs(votes.astype(int) - want).sum() / 2  # votes that changed class
    assert moved <= 0.001 * 12 * want[0].size, f"{moved} of {12 * want[0].size} votes differ"


# ---------------------------------------------------------------- random-init golden vectors
def test_golden_single_axis(predictor, golden):
    from volume_segmantics.utilities.base_data_utils import Axis

    labels, probs = predictor._predict_single_axis(golden["volume"], axis=Axis.Y)
    full = golden["full_probs_d1"]  # [S,C,H,W] slice space of direction 1 = (Y; Z, X)
    top2 = np.sort(full, axis=1)[:, -2:]
    margin = (top2[:, 1] - top2[:, 0]).swapaxes(0, 1)
    _check("random-init Y", labels, probs, golden["low_y_labels"], golden["low_y_probs"], RANDOM_INIT_FLOOR, margin)


def test_golden_three_and_twelve_ways(predictor, golden, unet_r34):
    oracle = po.OraclePredictor(unet_r34[0], 4)  # the weights the golden file was generated from
    vol = golden["volume"]
    labels, probs = predictor._predict_3_ways_max_probs(vol)
    _check("random-init 3-way", labels, probs, golden["medium_labels"], golden["medium_probs"], RANDOM_INIT_FLOOR,
           _merged_margin(oracle, vol, range(3)))
    labels, probs = predictor._predict_12_ways_max_probs(vol)
    _check("random-init 12-way", labels, probs, golden["high_labels"], golden["high_probs"], RANDOM_INIT_FLOOR,
           _merged_margin(oracle, vol, range(12)))


def test_manager_quality_dispatch(model_path, golden):
    from volume_segmantics.model import VolSeg2DPredictionManager
    from volume_segmantics.utilities import Quality

    mgr = VolSeg2DPredictionManager(str(model_path), golden["volume"].astype(np.int64), SimpleNamespace(**SETTINGS))
    out = mgr.predict_volume_to_path(None, Quality.MEDIUM)
    assert out.shape == golden["volume"].shape and out.dtype == np.uint8
    assert (out == golden["medium_labels"]).mean() >= RANDOM_INIT_FLOOR
    s = dict(SETTINGS, prediction_axis="y")
    mgr = VolSeg2DPredictionManager(str(model_path), golden["volume"], SimpleNamespace(**s))
    out = mgr.predict_volume_to_path(None, Quality.LOW)
    assert out.shape == golden["volume"].shape  # reference tests/test_vol_seg_prediction_manager.py:40-64
    assert (out == golden["low_y_labels"]).mean() >= RANDOM_INIT_FLOOR
    one_hot = VolSeg2DPredictionManager(str(model_path), golden["volume"], SimpleNamespace(**dict(SETTINGS, one_hot=True)))
    votes = one_hot.predict_volume_to_path(None, Quality.LOW)
    assert votes.dtype == np.uint8 and votes.ndim == 4 and (votes.sum(0) == 1).all()


def test_skip_duplicates_is_result_identical(engine, unet_r34, golden):
    _, model = unet_r34
    engine.load_model(model)
    engine.set_volume(golden["volume"])
    engine.predict((1 << 12) - 1, skip_duplicates=True)
    a = engine.fetch()
    engine.reset()
    engine.predict((1 << 12) - 1, skip_duplicates=False)
    b = engine.fetch()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_kernel_variants_agree(engine, unet_r34, golden):
    """Fused head == separate head kernel bit for bit (same logits, same arithmetic);
    halo kernels vs per-tap TMA kernel vs CUDA-core kernel agree to accumulation order."""
    _, model = unet_r34
    engine.load_model(model)
    engine.set_volume(golden["volume"])
    results = {}
    for name, flags in {"default": {}, "fused": {"fuse_head": 1}, "pertap": {"halo": 0}}.items():
        for k, v in flags.items():
            engine.set_flag(k, v)
        engine.reset()
        engine.predict(0b111, True)
        results[name] = engine.fetch()
        for k in flags:
            engine.set_flag(k, {"fuse_head": 0, "halo": 1}[k])
    assert np.array_equal(results["default"][0], results["fused"][0])
    assert np.array_equal(results["default"][1], results["fused"][1])
    dp = np.abs(results["default"][1].astype(np.float32) - results["pertap"][1].astype(np.float32)).max()
    assert dp < 2e-3 and (results["default"][0] == results["pertap"][0]).mean() > 0.99


def test_unsliceable_dtypes_are_refused_with_the_remedy(predictor):
    """float64 / float16 volumes fail in the reference too (double batch into float32 weights)."""
    with pytest.raises(TypeError, match="float32"):
        predictor._predict_single_axis(np.random.rand(8, 32, 32))
    with pytest.raises(ValueError, match="int32"):
        predictor._predict_single_axis(np.full((8, 32, 32), 2**40, np.int64))


def test_two_checkpoints_one_engine(tmp_path, unet_r34, trained_unet_r34, golden):
    """a second predictor on the same engine, and weights changed in place, must never
    run the previous plan (the plan key is the module object + a per-tensor version fingerprint)."""
    from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor

    vol = golden["volume"]
    pa = VolSeg2dPredictor(str(_save(unet_r34[0], tmp_path / "a.pytorch")), SimpleNamespace(**SETTINGS))
    pb = VolSeg2dPredictor(str(_save(trained_unet_r34[0], tmp_path / "b.pytorch")), SimpleNamespace(**SETTINGS))
    la, _ = pa._predict_single_axis(vol)
    lb, _ = pb._predict_single_axis(vol)
    la2, _ = pa._predict_single_axis(vol)
    assert np.array_equal(la, la2) and not np.array_equal(la, lb)
    pa.model.load_state_dict(trained_unet_r34[0].state_dict())  # in place, same module object
    la3, _ = pa._predict_single_axis(vol)
    assert np.array_equal(la3, lb)


def test_one_hot_class_count_change_on_one_shape(engine, golden):
    """the vote volume is sized by the class count of the loaded plan."""
    from volume_segmantics_b200.plan import B200SegmentationModel

    vol = golden["volume"]
    for classes in (2, 6, 3):
        torch.manual_seed(classes)
        model = B200SegmentationModel("U_NET", "resnet34", classes)
        engine.load_model(model)
        engine.set_volume(vol)
        engine.set_vote_mode(True)
        engine.reset()
        engine.predict(0b111, skip_duplicates=False)
        votes = engine.fetch_votes()
        engine.set_vote_mode(False)
        assert votes.shape == (classes,) + vol.shape and (votes.sum(0) == 3).all()
