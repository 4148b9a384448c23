"""Host logic of the space-to-depth tail rewrite (plan.py): the rewritten convolutions are the
same linear maps as smp's last DecoderBlock + segmentation head (nearest-x2 upsample ->
conv3x3 -> conv3x3 -> conv3x3), checked in fp32 with torch on the CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from volume_segmantics_b200.netspec import build_netspec
from volume_segmantics_b200.plan import find_s2d_tail, s2d_conv_weights, s2d_upconv_weights


def _unshuffle(y: torch.Tensor, c: int) -> torch.Tensor:
    """S2D [N, 4*c, H, W] (channel = (2a+b)*c + k) -> [N, c, 2H, 2W]."""
    n, _, h, w = y.shape
    y = y.reshape(n, 2, 2, c, h, w)  # a, b, k, i, j
    return y.permute(0, 3, 4, 1, 5, 2).reshape(n, c, 2 * h, 2 * w)


@pytest.mark.parametrize("h,w", [(4, 6), (1, 1), (5, 3)])
def test_s2d_chain_equals_upsample_conv_chain(h, w):
    g = torch.Generator().manual_seed(5)
    cin, ca, cb, classes = 32, 16, 16, 4
    x = torch.randn(2, cin, h, w, generator=g)
    wa, ba = torch.randn(ca, cin, 3, 3, generator=g), torch.randn(ca, generator=g)
    wb, bb = torch.randn(cb, ca, 3, 3, generator=g), torch.randn(cb, generator=g)
    wh, bh = torch.randn(classes, cb, 3, 3, generator=g), torch.randn(classes, generator=g)
    with torch.no_grad():
        up = F.interpolate(x, scale_factor=2, mode="nearest")
        ya = F.relu(F.conv2d(up, wa, ba, padding=1))
        yb = F.relu(F.conv2d(ya, wb, bb, padding=1))
        want = F.conv2d(yb, wh, bh, padding=1)
        sa = F.relu(F.conv2d(x, s2d_upconv_weights(wa), ba.repeat(4), padding=1))
        sb = F.relu(F.conv2d(sa, s2d_conv_weights(wb), bb.repeat(4), padding=1))
        got = F.conv2d(sb, s2d_conv_weights(wh), bh.repeat(4), padding=1)
    assert torch.allclose(_unshuffle(sa, ca), ya, atol=1e-4, rtol=1e-5)
    assert torch.allclose(_unshuffle(sb, cb), yb, atol=1e-3, rtol=1e-5)
    assert torch.allclose(_unshuffle(got, classes), want, atol=1e-2, rtol=1e-5)


def test_s2d_conv_weights_sparsity():
    """Exactly 16 of the 36 (tap, input sub-pixel) blocks are non-zero: the 4x4 patch of a 2x2 block."""
    w = torch.ones(16, 16, 3, 3)
    s = s2d_conv_weights(w).reshape(4, 16, 4, 16, 3, 3)
    used = (s.abs().sum(dim=(0, 1, 3)) > 0)  # [input sub-pixel, dy, dx]
    assert int(used.sum()) == 16
    # every (output sub-pixel, ky, kx) lands on exactly one block
    assert float(s.sum()) == 4 * 9 * 16 * 16


def test_find_s2d_tail_matches_unet_and_unetplusplus_only():
    for arch, enc in (("U_NET", "resnet34"), ("U_NET_PLUS_PLUS", "resnext50_32x4d"), ("U_NET", "resnet50")):
        spec = build_netspec(arch, enc, 4)
        tail = find_s2d_tail(spec)
        assert tail is not None, arch
        a, b, h, head = tail
        assert spec.layers[a].srcs[0][1] == 1 and spec.layers[a].cout == 16
        assert spec.layers[b].cin == 16 and spec.layers[h].name == "segmentation_head.0"
        assert head == len(spec.layers) - 1
    assert find_s2d_tail(build_netspec("DEEPLABV3_PLUS", "resnet50", 4)) is None
    assert find_s2d_tail(build_netspec("U_NET", "resnet34", 12)) is None  # > 8 classes: plain head path


def test_s2d_disabled_by_env(monkeypatch):
    monkeypatch.setenv("VSB200_S2D_TAIL", "0")
    assert find_s2d_tail(build_netspec("U_NET", "resnet34", 4)) is None
