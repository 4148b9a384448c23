"""Ingest either side of the uint8 fast path.

* Typed slicer: volumes the reference would slice as they are -- integers of any depth are cast
  to float32 and divided by 255, float32 is fed unchanged (datasets.py:129-135) -- bit-exact in
  the build's 16-bit format against the oracle's cv2 / numpy restatement, all 12 directions,
  ragged shapes (reflect-101 pads, pad > image).
* The same volumes end to end through ``VolSeg2dPredictor`` against the oracle.
* ``BaseDataManager._preprocess_data`` on the GPU (base_data_manager.py:29-42,
  base_data_utils.py:243-287): statistics against numpy (tolerance stated per dtype), the
  clip / rescale / quantise pass bit-exact given the same statistics (float32 arithmetic for
  float32 data), the clipped-voxel counts exact, and the result left resident in HBM.
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import act_bits
from oracle import predict_oracle as po

pytestmark = pytest.mark.gpu

PROB_TOL = 2e-2
SETTINGS = dict(quality="medium", output_probs=False, clip_data=False, st_dev_factor=2.575,
                data_hdf5_path="/data", cuda_device=0, downsample=False, one_hot=False, prediction_axis="Z")


def _typed_volume(dtype, shape, seed):
    rng = np.random.default_rng(seed)
    if dtype == np.float32:
        return rng.normal(0.45, 0.25, shape).astype(np.float32)  # already "normalised" intensities
    if dtype == np.uint8:
        return rng.integers(0, 256, shape, dtype=np.uint8)
    if dtype == np.int8:
        return rng.integers(-128, 128, shape, dtype=np.int8)
    if dtype == np.uint16:
        return rng.integers(0, 65536, shape, dtype=np.uint16)
    if dtype == np.int16:
        return rng.integers(-32768, 32768, shape, dtype=np.int16)
    return rng.integers(-1_000_000, 1_000_000, shape, dtype=np.int32)


@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.int32, np.int8, np.float32, np.uint8])
@pytest.mark.parametrize("shape", [(9, 61, 33), (33, 29, 70), (40, 10, 13)])
def test_typed_slicer_bit_exact_all_directions(engine, dtype, shape):
    vol = _typed_volume(dtype, shape, 17)
    engine.set_volume(vol)
    for d in range(12):
        g = engine.geometry(d)
        got = engine.slice_batch(d, 0, g.S, generic=True)
        want = act_bits(po.slicer_oracle(vol, d))
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"{np.dtype(dtype).name} direction {d}: {(got != want).sum()} differing pixels"


def test_generic_slicer_equals_fast_uint8_slicer(engine):
    vol = _typed_volume(np.uint8, (37, 45, 70), 3)
    engine.set_volume(vol)
    for d in range(12):
        g = engine.geometry(d)
        assert np.array_equal(engine.slice_batch(d, 0, g.S), engine.slice_batch(d, 0, g.S, generic=True))
        assert np.array_equal(engine.slice_batch(d, 2, g.S - 3), engine.slice_batch(d, 2, g.S - 3, generic=True))


@pytest.fixture(scope="module")
def trained_pair(tmp_path_factory, trained_unet_r34):
    import volume_segmantics.utilities.base_data_utils as utils
    from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor

    oracle, _ = trained_unet_r34
    path = tmp_path_factory.mktemp("ingest") / "m.pytorch"
    struc = {"type": utils.ModelType.U_NET, "encoder_name": "resnet34", "encoder_weights": None,
             "in_channels": 1, "classes": 4}
    torch.save({"model_state_dict": oracle.state_dict(), "model_struc_dict": struc, "label_codes": {}}, path)
    return path, VolSeg2dPredictor(str(path), SimpleNamespace(**SETTINGS)), po.OraclePredictor(oracle, 4)


@pytest.mark.parametrize("dtype", [np.uint16, np.float32, np.int64])
def test_non_uint8_volumes_end_to_end(trained_pair, dtype):
    """clip_data: False with 16-bit or float data works in the reference (datasets.py:129-135)."""
    _, pred, oracle = trained_pair
    shape = (12, 40, 45)
    rng = np.random.default_rng(4)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    base = np.clip(128 + 50 * np.sin(z / 4.0) + 40 * np.sin(y / 6.0 + 1) + 30 * np.sin(x / 8.0 + 2)
                   + rng.normal(0, 20, shape), 0, 255)
    if dtype == np.float32:
        vol = (base / 255).astype(np.float32)  # what a float volume must look like for the /255-free branch
    else:
        vol = (base * 1.7).astype(dtype)       # integers beyond 255: still divided by 255, as the reference does
    labels, probs = pred._predict_3_ways_max_probs(vol)
    want_l, want_p = oracle.predict_3_ways_max_probs(vol)
    perr = np.abs(probs.astype(np.float32) - want_p.astype(np.float32)).max()
    bad = labels != want_l
    cb = np.sort(oracle.class_best_over_directions(vol, range(3)), axis=0)
    worst = (cb[-1] - cb[-2])[bad].max() if bad.any() else 0.0
    print(f"[ingest {np.dtype(dtype).name}] agreement {1 - bad.mean():.5f} max prob err {perr:.5f}")
    assert perr < PROB_TOL and worst < PROB_TOL and 1 - bad.mean() >= 0.999


# --------------------------------------------------------------------------- f-1 on the GPU
def _raw_volume(dtype, seed, with_nan):
    rng = np.random.default_rng(seed)
    vol = rng.normal(500, 200, (37, 64, 129))
    if np.issubdtype(dtype, np.floating):
        vol = vol.astype(dtype)
        if with_nan:
            vol[rng.integers(0, 37, 60), rng.integers(0, 64, 60), rng.integers(0, 129, 60)] = np.nan
    else:
        vol = vol.clip(0, None).astype(dtype)
    return vol


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint16, np.int32, np.uint8])
def test_gpu_moments_against_numpy(engine, dtype):
    """float64 / integer data: numpy accumulates in float64 too -> agreement to rounding (1e-12).
    float32 data: numpy's pairwise float32 sums carry ~1e-7 relative error themselves; the GPU's
    float64 reduction, rounded to float32 as numpy returns it, agrees to a few float32 ulps."""
    vol = _raw_volume(dtype, 6, with_nan=True)
    engine.raw_upload(vol)
    count, mean, std, nans = engine.raw_moments()
    count2, mean2, std2, nans2 = engine.raw_moments()  # fixed reduction order: identical bits
    engine.raw_release()
    assert (count, mean, std, nans) == (count2, mean2, std2, nans2)
    assert nans == int(np.isnan(vol).sum()) and count == vol.size - nans
    rtol = 3e-6 if dtype == np.float32 else 1e-12
    assert abs(mean - float(np.nanmean(vol))) <= rtol * abs(float(np.nanmean(vol)))
    assert abs(std - float(np.nanstd(vol))) <= rtol * float(np.nanstd(vol))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint16, np.int32, np.int64, np.uint8])
def test_gpu_clip_bit_exact_given_statistics(engine, dtype):
    """Both the host-buffer entry (vsb_clip_to_uint8) and the resident one (vsb_raw_clip_to_volume):
    identical to numpy when handed numpy's own mean / bounds -- float32 data in float32 arithmetic."""
    vol = np.concatenate([_raw_volume(dtype, 2, True)] * 7)  # 2.1 M voxels: float32-vs-float64 arithmetic would show
    mean = np.nanmean(vol)
    want = po.clip_to_uint8_oracle(vol, mean, 2.575)
    sd = np.nanstd(vol)
    lower, upper = mean - sd * 2.575, mean + sd * 2.575
    got = engine.clip_to_uint8(vol, float(mean), float(lower), float(upper))
    assert np.array_equal(got, want), f"host-buffer clip: {(got != want).sum()} voxels differ"
    engine.raw_upload(vol)
    out, gt, lt = engine.raw_clip_to_volume(float(mean), float(lower), float(upper), vol.shape)
    assert np.array_equal(out, want), f"resident clip: {(out != want).sum()} voxels differ"
    with np.errstate(invalid="ignore"):
        assert gt == int((vol > upper).sum()) and lt == int((vol < lower).sum())
    assert not out.flags.writeable and engine.holds(out) and not engine.holds(out.copy())


def test_manager_preprocess_on_gpu_and_volume_stays_resident(trained_pair):
    from volume_segmantics.model import VolSeg2DPredictionManager
    from volume_segmantics.utilities import Quality

    path, _, oracle = trained_pair
    rng = np.random.default_rng(8)
    shape = (20, 40, 45)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    raw = (3000 + 900 * np.sin(z / 4.0) + 700 * np.sin(y / 6.0 + 1) + 500 * np.sin(x / 8.0 + 2)
           + rng.normal(0, 300, shape)).astype(np.float32)
    raw[3, 5, 7] = np.nan
    settings = SimpleNamespace(**dict(SETTINGS, clip_data=True))
    mgr = VolSeg2DPredictionManager(str(path), raw.copy(), settings)
    want_u8 = po.clip_to_uint8_oracle(raw, np.nanmean(raw), 2.575)
    assert mgr.data_vol.dtype == np.uint8 and mgr.data_vol.shape == shape
    # statistics agree with numpy to float32 rounding, so at most a handful of voxels may sit one grey level off
    diff = np.abs(mgr.data_vol.astype(int) - want_u8.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    assert abs(float(mgr.data_mean) - float(np.nanmean(raw))) < 1e-5 * abs(float(np.nanmean(raw)))
    eng = mgr.predictor.engine
    assert eng.holds(mgr.data_vol)  # nothing to upload for the prediction
    gen = eng.volume_generation()
    labels = mgr.predict_volume_to_path(None, Quality.MEDIUM)
    assert eng.volume_generation() == gen, "the resident volume was uploaded again"
    again = mgr.predictor._predict_3_ways_max_probs(np.array(mgr.data_vol))[0]  # a copy: goes through the upload
    assert np.array_equal(labels, again)
    want_l, _ = oracle.predict_3_ways_max_probs(np.array(mgr.data_vol))
    assert (labels == want_l).mean() >= 0.999
