"""Slicer kernel (a): bit-exact (in the build's 16-bit format) against the oracle's cv2/numpy restatement of
datasets.py:120-142 + augmentations.py:46-65, for all 12 directions, dims that
exercise pad = 0,1,2,3 (mod 4) and pads larger than the image."""
import hashlib
import json

import numpy as np
import pytest

from conftest import act_bits, act_tag
from oracle import make_golden as mg
from oracle import predict_oracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("si", range(len(mg.SLICER_SHAPES)))
def test_slicer_bit_exact_all_directions(engine, golden_dir, si):
    digests = json.loads((golden_dir / "slicer_digests.json").read_text())
    shape = mg.SLICER_SHAPES[si]
    vol = mg.synth_volume(shape, 100 + si)
    engine.set_volume(vol)
    for d in range(12):
        g = engine.geometry(d)
        got = engine.slice_batch(d, 0, g.S)
        want = act_bits(po.slicer_oracle(vol, d))
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"direction {d}: {(got != want).sum()} differing pixels"
        assert hashlib.sha256(got.tobytes()).hexdigest() == digests[f"{shape}|{d}"]["sha256_" + act_tag()]


def test_slicer_small_golden_file(engine, golden_dir):
    z = np.load(golden_dir / "slicer_small.npz")
    vol = mg.synth_volume(mg.SLICER_SHAPES[0], 100)
    engine.set_volume(vol)
    for d in range(12):
        assert np.array_equal(engine.slice_batch(d, 0, engine.geometry(d).S), act_bits(z[f"d{d}"]))


def test_slicer_ragged_batches_and_offsets(engine):
    vol = mg.synth_volume((40, 37, 70), 5)
    engine.set_volume(vol)
    for d in (0, 2, 5, 7, 11):
        g = engine.geometry(d)
        want = act_bits(po.slicer_oracle(vol, d))
        for s0, nb in [(0, 1), (3, 9), (g.S - 33 if g.S > 33 else 0, min(33, g.S)), (g.S - 1, 1)]:
            assert np.array_equal(engine.slice_batch(d, s0, nb), want[s0:s0 + nb])


def test_slicer_full_size_properties(engine):
    """512^3 (BASELINE config 2 size): padding is the identity, so the slicer is a
    pure permutation + LUT: per-direction histograms of the bf16 output equal the
    LUT-mapped histogram of the volume slab."""
    rng = np.random.default_rng(20240)
    vol = rng.integers(0, 256, size=(512, 512, 512), dtype=np.uint8)
    engine.set_volume(vol)
    lut = act_bits(((np.arange(256, dtype=np.float32) / 255) - np.float32(0.449)) / np.float32(0.226))
    for d in (0, 4, 5, 8):
        got = engine.slice_batch(d, 100, 32)
        ref = lut[po.direction_slices(vol, d)[100:132]]
        assert np.array_equal(got, ref)
