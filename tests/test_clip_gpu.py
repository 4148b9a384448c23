"""clip_to_uint8 (base_data_utils.py:243-287) through the drop-in function: statistics
from numpy as in the reference, the elementwise passes as one GPU kernel in numpy's own precision
(float32 for float32 data, float64 otherwise).  Bit-exact.  The all-GPU pre-processing of
BaseDataManager is covered by tests/test_ingest_gpu.py."""
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import predict_oracle as po


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint16, np.int64, np.uint8])
def test_host_clip_matches_oracle(dtype):
    """CPU: the numpy path of the drop-in equals the restated reference function."""
    from volume_segmantics.utilities.base_data_utils import clip_to_uint8

    rng = np.random.default_rng(1)
    vol = (rng.normal(1000, 300, (9, 20, 31)) if np.issubdtype(dtype, np.integer) and dtype != np.uint8
           else rng.normal(100, 40, (9, 20, 31))).clip(0, None).astype(dtype)
    mean = np.nanmean(vol)
    want = po.clip_to_uint8_oracle(vol, mean, 2.575)
    got = clip_to_uint8(vol.copy(), mean, 2.575)
    assert got.dtype == np.uint8 and np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint16, np.int32, np.int64])
def test_gpu_clip_bit_exact(engine, dtype):
    from volume_segmantics.utilities.base_data_utils import clip_to_uint8

    rng = np.random.default_rng(2)
    vol = rng.normal(500, 200, (37, 64, 129))
    if np.issubdtype(dtype, np.floating):
        vol = vol.astype(dtype)
        vol[rng.integers(0, 37, 50), rng.integers(0, 64, 50), rng.integers(0, 129, 50)] = np.nan
    else:
        vol = vol.clip(0, None).astype(dtype)
    mean = np.nanmean(vol)
    want = po.clip_to_uint8_oracle(vol, mean, 2.575)
    got = clip_to_uint8(vol.copy(), mean, 2.575, cuda_device=0)
    assert got.dtype == np.uint8 and got.shape == vol.shape
    assert np.array_equal(got, want), f"{(got != want).sum()} voxels differ"


@pytest.mark.gpu
def test_manager_preprocess_uses_gpu_clip_and_matches(engine):
    from volume_segmantics.data.base_data_manager import BaseDataManager

    vol = np.random.default_rng(3).normal(0, 1, (20, 33, 45)).astype(np.float32)
    settings = SimpleNamespace(st_dev_factor=2.575, downsample=False, clip_data=True, cuda_device=0,
                               data_hdf5_path="/data")
    mgr = BaseDataManager(vol.copy(), settings)
    want = po.clip_to_uint8_oracle(vol, np.nanmean(vol), 2.575)
    # statistics come from the GPU's float64 reduction (numpy: float32 pairwise sums): bounds agree to a
    # few float32 ulps, so a voxel sitting exactly on a quantisation step may land one grey level off
    diff = np.abs(mgr.data_vol.astype(int) - want.astype(int))
    assert mgr.data_vol.dtype == np.uint8 and diff.max() <= 1 and (diff != 0).mean() < 1e-3
