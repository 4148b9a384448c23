"""Oracle restatement of the segmentation_models_pytorch (^0.2.1) networks the
reference builds in ``volume_segmantics/model/model_2d.py:10-39``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  fp32, CPU, plain torch.

smp itself is not installed here, so the three architectures named by
BASELINE.json are restated from the package's published structure, on top of
the *same* ``torchvision.models.resnet.ResNet`` class that smp's
``ResNetEncoder`` subclasses.  Parameter names equal smp's state-dict keys so a
reference-format ``.pytorch`` file loads with ``strict=True``:

  encoder.conv1 / bn1 / layer{1..4}.{i}.{conv,bn}{1,2,3} / downsample.{0,1}
  decoder.blocks.{i}.conv{1,2}.{0,1}                      (Unet)
  decoder.blocks.x_{d}_{l}.conv{1,2}.{0,1}                (UnetPlusPlus)
  decoder.aspp.0.convs.{0..4}.*, decoder.aspp.0.project.*, decoder.aspp.{1,2},
  decoder.block1.{0,1}, decoder.block2.{0,1}              (DeepLabV3Plus)
  segmentation_head.0.{weight,bias}
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet

# encoder_name -> (block, layers, groups, width_per_group, feature channels)
ENCODERS = {
    "resnet18": (BasicBlock, [2, 2, 2, 2], 1, 64, (64, 64, 128, 256, 512)),
    "resnet34": (BasicBlock, [3, 4, 6, 3], 1, 64, (64, 64, 128, 256, 512)),
    "resnet50": (Bottleneck, [3, 4, 6, 3], 1, 64, (64, 256, 512, 1024, 2048)),
    "resnet101": (Bottleneck, [3, 4, 23, 3], 1, 64, (64, 256, 512, 1024, 2048)),
    "resnext50_32x4d": (Bottleneck, [3, 4, 6, 3], 32, 4, (64, 256, 512, 1024, 2048)),
}


class OracleEncoder(ResNet):
    """torchvision ResNet trunk returning the 5 feature maps smp's
    ``ResNetEncoder.forward`` returns after the identity stage."""

    def __init__(self, name: str, in_channels: int = 1):
        block, layers, groups, width, chans = ENCODERS[name]
        super().__init__(block, layers, groups=groups, width_per_group=width)
        del self.fc
        del self.avgpool
        self.feature_channels = chans
        if in_channels != 3:
            self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)

    def _dilate(self, layer, rate):
        # smp `replace_strides_with_dilation(layer, rate)`: EVERY Conv2d of the stage
        for mod in layer.modules():
            if isinstance(mod, nn.Conv2d):
                mod.stride = (1, 1)
                mod.dilation = (rate, rate)
                kh, _ = mod.kernel_size
                mod.padding = ((kh // 2) * rate, (kh // 2) * rate)

    def make_dilated_os16(self):
        self._dilate(self.layer4, 2)

    def make_dilated_os8(self):
        # smp DeepLabV3: encoder.make_dilated(stage_list=[4, 5], dilation_list=[2, 4])
        self._dilate(self.layer3, 2)
        self._dilate(self.layer4, 4)

    def forward(self, x):
        f1 = self.relu(self.bn1(self.conv1(x)))
        f2 = self.layer1(self.maxpool(f1))
        f3 = self.layer2(f2)
        f4 = self.layer3(f3)
        f5 = self.layer4(f4)
        return [f1, f2, f3, f4, f5]


def _conv_bn_relu(cin, cout, k, padding=0):
    # smp Conv2dReLU(nn.Sequential): "0" conv (no bias), "1" bn, "2" relu
    return nn.Sequential(
        nn.Conv2d(cin, cout, k, padding=padding, bias=False),
        nn.BatchNorm2d(cout),
        nn.ReLU(inplace=True),
    )


class OracleDecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _conv_bn_relu(cin + cskip, cout, 3, 1)
        self.conv2 = _conv_bn_relu(cout, cout, 3, 1)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class _UnetDecoder(nn.Module):
    def __init__(self, enc_ch, dec_ch=(256, 128, 64, 32, 16)):
        super().__init__()
        rev = list(enc_ch[::-1])
        ins = [rev[0]] + list(dec_ch[:-1])
        skips = rev[1:] + [0]
        self.blocks = nn.ModuleList(
            OracleDecoderBlock(i, s, o) for i, s, o in zip(ins, skips, dec_ch)
        )

    def forward(self, feats):
        rev = feats[::-1]
        x = rev[0]
        for i, blk in enumerate(self.blocks):
            x = blk(x, rev[i + 1] if i + 1 < len(rev) else None)
        return x


class _UnetPPDecoder(nn.Module):
    def __init__(self, enc_ch, dec_ch=(256, 128, 64, 32, 16)):
        super().__init__()
        rev = list(enc_ch[::-1])
        self.ins = [rev[0]] + list(dec_ch[:-1])
        self.skips = rev[1:] + [0]
        self.outs = list(dec_ch)
        blocks = {}
        n = len(self.ins)
        for l in range(n - 1):
            for d in range(l + 1):
                if d == 0:
                    ci, cs, co = self.ins[l], self.skips[l] * (l + 1), self.outs[l]
                else:
                    co = self.skips[l]
                    cs = self.skips[l] * (l + 1 - d)
                    ci = self.skips[l - 1]
                blocks[f"x_{d}_{l}"] = OracleDecoderBlock(ci, cs, co)
        blocks[f"x_0_{n - 1}"] = OracleDecoderBlock(self.ins[-1], 0, self.outs[-1])
        self.blocks = nn.ModuleDict(blocks)
        self.depth = n - 1

    def forward(self, feats):
        f = feats[::-1]
        dense = {}
        for l in range(len(self.ins) - 1):
            for d in range(self.depth - l):
                if l == 0:
                    dense[f"x_{d}_{d}"] = self.blocks[f"x_{d}_{d}"](f[d], f[d + 1])
                else:
                    li = d + l
                    cat = [dense[f"x_{i}_{li}"] for i in range(d + 1, li + 1)]
                    cat = torch.cat(cat + [f[li + 1]], dim=1)
                    dense[f"x_{d}_{li}"] = self.blocks[f"x_{d}_{li}"](
                        dense[f"x_{d}_{li - 1}"], cat
                    )
        dp = self.depth
        dense[f"x_0_{dp}"] = self.blocks[f"x_0_{dp}"](dense[f"x_0_{dp - 1}"])
        return dense[f"x_0_{dp}"]


def _separable(cin, cout, k, padding, dilation=1):
    # smp SeparableConv2d(nn.Sequential): "0" depthwise, "1" pointwise, no bias
    return nn.Sequential(
        nn.Conv2d(cin, cin, k, padding=padding, dilation=dilation, groups=cin, bias=False),
        nn.Conv2d(cin, cout, 1, bias=False),
    )


class _ASPPPool(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(
            nn.AdaptiveAvgPool2d(1),
            nn.Conv2d(cin, cout, 1, bias=False),
            nn.BatchNorm2d(cout),
            nn.ReLU(),
        )

    def forward(self, x):
        size = x.shape[-2:]
        for m in self:
            x = m(x)
        return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


class _ASPP(nn.Module):
    def __init__(self, cin, cout, rates, separable=True):
        super().__init__()
        mods = [nn.Sequential(nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())]
        for r in rates:
            # smp ASPPSeparableConv (DeepLabV3+) / ASPPConv (DeepLabV3): conv, BN, ReLU
            conv = _separable(cin, cout, 3, r, r) if separable else nn.Conv2d(cin, cout, 3, padding=r, dilation=r, bias=False)
            mods.append(nn.Sequential(conv, nn.BatchNorm2d(cout), nn.ReLU()))
        mods.append(_ASPPPool(cin, cout))
        self.convs = nn.ModuleList(mods)
        self.project = nn.Sequential(
            nn.Conv2d(5 * cout, cout, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(), nn.Dropout(0.5)
        )

    def forward(self, x):
        return self.project(torch.cat([c(x) for c in self.convs], dim=1))


class _DeepLabV3PlusDecoder(nn.Module):
    def __init__(self, enc_ch, out_ch=256, rates=(12, 24, 36)):
        super().__init__()
        self.aspp = nn.Sequential(
            _ASPP(enc_ch[-1], out_ch, rates),
            _separable(out_ch, out_ch, 3, 1),
            nn.BatchNorm2d(out_ch),
            nn.ReLU(),
        )
        self.up = nn.UpsamplingBilinear2d(scale_factor=4)  # align_corners=True
        hi_in, hi_out = enc_ch[-4], 48
        self.block1 = nn.Sequential(
            nn.Conv2d(hi_in, hi_out, 1, bias=False), nn.BatchNorm2d(hi_out), nn.ReLU()
        )
        self.block2 = nn.Sequential(
            _separable(hi_out + out_ch, out_ch, 3, 1), nn.BatchNorm2d(out_ch), nn.ReLU()
        )

    def forward(self, feats):
        a = self.up(self.aspp(feats[-1]))
        h = self.block1(feats[-4])
        return self.block2(torch.cat([a, h], dim=1))


class _DeepLabV3Decoder(nn.Sequential):
    """smp DeepLabV3Decoder(nn.Sequential) [ext]: ASPP (plain atrous 3x3), conv3x3, BN, ReLU on the last feature."""

    def __init__(self, cin, out_ch=256, rates=(12, 24, 36)):
        super().__init__(
            _ASPP(cin, out_ch, rates, separable=False),
            nn.Conv2d(out_ch, out_ch, 3, padding=1, bias=False),
            nn.BatchNorm2d(out_ch),
            nn.ReLU(),
        )

    def forward(self, feats):
        return super().forward(feats[-1])


class OracleSegModel(nn.Module):
    """``arch`` in {"unet", "unetplusplus", "deeplabv3plus", "deeplabv3"}."""

    def __init__(self, arch: str, encoder_name: str, classes: int, in_channels: int = 1):
        super().__init__()
        self.arch = arch
        self.encoder = OracleEncoder(encoder_name, in_channels)
        ch = self.encoder.feature_channels
        if arch == "unet":
            self.decoder = _UnetDecoder(ch)
            self.segmentation_head = nn.Sequential(nn.Conv2d(16, classes, 3, padding=1))
        elif arch == "unetplusplus":
            self.decoder = _UnetPPDecoder(ch)
            self.segmentation_head = nn.Sequential(nn.Conv2d(16, classes, 3, padding=1))
        elif arch == "deeplabv3plus":
            self.encoder.make_dilated_os16()
            self.decoder = _DeepLabV3PlusDecoder(ch)
            self.segmentation_head = nn.Sequential(
                nn.Conv2d(256, classes, 1), nn.UpsamplingBilinear2d(scale_factor=4)
            )
        elif arch == "deeplabv3":
            self.encoder.make_dilated_os8()
            self.decoder = _DeepLabV3Decoder(ch[-1])
            self.segmentation_head = nn.Sequential(
                nn.Conv2d(256, classes, 1), nn.UpsamplingBilinear2d(scale_factor=8)
            )
        else:
            raise ValueError(arch)

    def forward(self, x):
        return self.segmentation_head(self.decoder(self.encoder(x)))


# name of the reference's ModelType enum member -> oracle arch
ARCH_OF_MODELTYPE = {
    "U_NET": "unet",
    "U_NET_PLUS_PLUS": "unetplusplus",
    "DEEPLABV3_PLUS": "deeplabv3plus",
    "DEEPLABV3": "deeplabv3",
}


def randomise_bn_(model: nn.Module, gen: torch.Generator) -> None:
    """SURVEY.md section 8d synthetic weights: default conv init, then BN
    weight~U(.5,1.5), bias~N(0,.1), running_mean~N(0,.1), running_var~U(.5,1.5)
    so that BN folding is actually exercised."""
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)
                m.running_mean.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)
                m.running_var.copy_(torch.rand(m.bias.shape, generator=gen) + 0.5)


def make_random_model(arch: str, encoder: str, classes: int, seed: int = 0) -> OracleSegModel:
    torch.manual_seed(seed)
    model = OracleSegModel(arch, encoder, classes)
    g = torch.Generator().manual_seed(seed + 1)
    randomise_bn_(model, g)
    return model.eval()
