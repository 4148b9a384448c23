"""Direction geometry, padding and crop rules: C host arithmetic (libvsb200) and
the Python sharding helpers against numpy / cv2 / torchvision -- the very calls
the reference makes (vol_seg_2d_predictor.py:34,108; base_data_utils.py:125-138;
augmentations.py:30-65)."""
import numpy as np
import pytest
import torch
import torchvision.transforms.functional as TVF

from oracle import predict_oracle as po
from volume_segmantics_b200 import _lib
from volume_segmantics_b200.host.utils import get_padded_dimension
from volume_segmantics_b200.sharding import direction_dims

SHAPES = [(3, 4, 5), (7, 29, 30), (10, 61, 33), (32, 32, 32), (1, 9, 2)]


def test_padded_dimension_known_answers():
    # reference tests/test_augmentations.py:6-10
    for n, want in [(32, 32), (64, 64), (33, 64), (13, 32), (0, 0)]:
        assert get_padded_dimension(n) == want
        assert po.get_padded_dimension(n) == want


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("d", range(12))
def test_direction_address_map_matches_numpy(shape, d):
    coords = np.arange(np.prod(shape), dtype=np.int64).reshape(shape)
    sl = po.direction_slices(coords, d)  # np.rot90 + swapaxes views, as the reference
    g = _lib.direction_geometry(*shape, d)
    assert (g.S, g.H, g.W) == sl.shape == direction_dims(shape, d)
    s, r, c = np.meshgrid(np.arange(g.S), np.arange(g.H), np.arange(g.W), indexing="ij")
    addr = g.base + s * g.stride_s + r * g.stride_r + c * g.stride_c
    assert np.array_equal(addr, sl)
    # inverse mapping puts every slice-space value back on its own voxel
    assert np.array_equal(po.direction_to_volume(sl, d), coords)


@pytest.mark.parametrize("n", [10, 25, 27, 29, 30, 31, 32, 33, 61, 64, 100, 255])
def test_pad_and_crop_offsets(n):
    g = _lib.direction_geometry(2, n, n, 0)
    hp = po.get_padded_dimension(n)
    assert g.Hp == hp and g.Wp == hp
    assert g.pad_top == int((hp - n) / 2.0)
    ramp = torch.arange(hp)[None, :, None].expand(1, hp, hp)
    crop = TVF.center_crop(ramp, [n, n])
    assert int(crop[0, 0, 0]) == g.crop_top == po.crop_offsets(n, n)[0]
    # p == 3 (mod 4) is the banker's-rounding one-pixel shift
    assert (g.crop_top - g.pad_top == 1) == ((hp - n) % 4 == 3)


def test_duplicate_directions_are_identical_image_sets():
    vol = np.random.default_rng(0).integers(0, 256, (5, 6, 7), dtype=np.uint8)
    for dup, first in {3: 1, 6: 4, 9: 7, 10: 0}.items():
        a, b = po.direction_slices(vol, dup), po.direction_slices(vol, first)
        assert a.shape == b.shape and np.array_equal(a, b[::-1])
