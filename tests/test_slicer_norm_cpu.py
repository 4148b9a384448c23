"""CPU restatement of the check the engine runs on the device when it is created
(vsb_slicer_norm_selfcheck): the slicer's single fused multiply-add per pixel must round to the
same 16-bit value as the reference's fp32 formula (datasets.py:129-135, config.py:41-42:
x / 255, - 0.449, / 0.226) for every one of the 256 possible inputs, in both 16-bit formats."""
import numpy as np
import pytest
import torch

from oracle import predict_oracle as po

A_BITS, B_BITS = 0x3C8E25F0, 0xBFFE4D07  # VSB_NORM_A / VSB_NORM_B in csrc/kernels_simple.cu


def _const(bits):
    return np.array([bits], np.uint32).view(np.float32)[0]


def test_constants_are_the_rounded_exact_values():
    assert _const(A_BITS) == np.float32(1.0 / (255.0 * 0.226))
    assert _const(B_BITS) == np.float32(-0.449 / 0.226)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fma_normalisation_equals_reference_formula(dtype):
    v = np.arange(256, dtype=np.float32)
    ref = ((v / np.float32(255.0)) - np.float32(0.449)) / np.float32(0.226)  # numpy fp32, as the reference
    # an exact fma: the fp32 x fp32 product and the sum are exact in float64, then one rounding to fp32
    fma = (v.astype(np.float64) * np.float64(_const(A_BITS)) + np.float64(_const(B_BITS))).astype(np.float32)
    a = torch.from_numpy(ref).to(dtype).view(torch.int16).numpy()
    b = torch.from_numpy(fma).to(dtype).view(torch.int16).numpy()
    assert np.array_equal(a, b)
    assert np.abs(ref.view(np.int32) - fma.view(np.int32)).max() < 128  # a few dozen fp32 ulps apart at most


def test_reference_formula_is_what_the_oracle_applies():
    img = np.arange(256, dtype=np.uint8).reshape(16, 16)
    got = po.preprocess_slice(np.pad(img, ((8, 8), (8, 8)), mode="reflect"))[8:24, 8:24]
    v = img.astype(np.float32)
    want = ((v / np.float32(255.0)) - np.float32(0.449)) / np.float32(0.226)
    assert np.array_equal(got.astype(np.float32), want)
