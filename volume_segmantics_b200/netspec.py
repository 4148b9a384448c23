"""Network description for the B200 engine.

The reference builds its models with segmentation_models_pytorch
(``volume_segmantics/model/model_2d.py:10-39``).  Here each architecture is a
flat list of layer records over numbered activation tensors.  The same list
(1) names every parameter with smp's state-dict key, so a reference ``.pytorch``
file loads unchanged, and (2) is lowered -- after BatchNorm folding -- into the
``vsb_op`` table that libvsb200 executes (``plan.py``).

Supported: Unet / UnetPlusPlus / DeepLabV3Plus (BASELINE.json configs 1-5) and DeepLabV3 decoders on
ResNet-18/34/50/101 and ResNeXt-50_32x4d encoders.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# encoder name -> (block kind, blocks per stage, groups, width_per_group)
ENCODER_CFG = {
    "resnet18": ("basic", (2, 2, 2, 2), 1, 64),
    "resnet34": ("basic", (3, 4, 6, 3), 1, 64),
    "resnet50": ("bottleneck", (3, 4, 6, 3), 1, 64),
    "resnet101": ("bottleneck", (3, 4, 23, 3), 1, 64),
    "resnext50_32x4d": ("bottleneck", (3, 4, 6, 3), 32, 4),
}
DECODER_CHANNELS = (256, 128, 64, 32, 16)


@dataclass
class Layer:
    kind: str  # conv | maxpool | gap | upsample | head
    out: int = -1
    srcs: List[Tuple[int, int]] = field(default_factory=list)  # (tensor id, nearest-x2 flag)
    res: int = -1
    # conv
    name: str = ""  # state-dict prefix of the conv ("<name>.weight")
    bn: Optional[str] = None  # state-dict prefix of the BatchNorm folded into it
    has_bias: bool = False
    cin: int = 0
    cout: int = 0
    k: int = 1
    stride: int = 1
    pad: int = 0
    dil: int = 1
    groups: int = 1
    relu: bool = False
    init: str = "encoder"  # which smp/torchvision initialiser applies
    # upsample / head
    mode: int = 0
    factor: int = 1


@dataclass
class TensorSpec:
    channels: int
    ds_log2: int  # spatial = padded >> ds_log2 ; -1 = 1x1
    dtype: int = 0  # 0 bf16, 1 f32


class NetSpec:
    def __init__(self, arch: str, encoder: str, classes: int, in_channels: int = 1):
        self.arch, self.encoder, self.classes, self.in_channels = arch, encoder, classes, in_channels
        self.tensors: List[TensorSpec] = [TensorSpec(in_channels, 0)]
        self.layers: List[Layer] = []

    # -- builders ---------------------------------------------------------------
    def tensor(self, channels: int, ds_log2: int, dtype: int = 0) -> int:
        self.tensors.append(TensorSpec(channels, ds_log2, dtype))
        return len(self.tensors) - 1

    def conv(self, name, srcs, cout, k, stride=1, pad=0, dil=1, groups=1, bn=None, bias=False,
             relu=False, res=-1, init="encoder", out_dtype=0) -> int:
        srcs = [(s, 0) if isinstance(s, int) else s for s in srcs]
        cin = sum(self.tensors[t].channels for t, _ in srcs)
        t0, up0 = srcs[0]
        ds = self.tensors[t0].ds_log2 - (1 if up0 else 0)
        if ds >= 0 and stride == 2:
            ds += 1
        out = self.tensor(cout, ds, out_dtype)
        self.layers.append(Layer("conv", out, srcs, res, name, bn, bias, cin, cout, k, stride, pad,
                                 dil, groups, relu, init))
        return out

    def maxpool(self, src) -> int:
        t = self.tensors[src]
        out = self.tensor(t.channels, t.ds_log2 + 1)
        self.layers.append(Layer("maxpool", out, [(src, 0)]))
        return out

    def gap(self, src) -> int:
        out = self.tensor(self.tensors[src].channels, -1)
        self.layers.append(Layer("gap", out, [(src, 0)]))
        return out

    def upsample(self, src, ds_log2, mode, factor=1) -> int:
        out = self.tensor(self.tensors[src].channels, ds_log2)
        self.layers.append(Layer("upsample", out, [(src, 0)], mode=mode, factor=factor))
        return out

    def head(self, logits, factor=1) -> None:
        self.layers.append(Layer("head", -1, [(logits, 0)], factor=factor))

    # -- parameter inventory (smp / torchvision key names) ------------------------
    def param_shapes(self):
        """Ordered {key: (shape, role)} for every parameter and buffer."""
        out = {}
        for L in self.layers:
            if L.kind != "conv":
                continue
            out[f"{L.name}.weight"] = ((L.cout, L.cin // L.groups, L.k, L.k), ("conv_w", L.init))
            if L.has_bias:
                out[f"{L.name}.bias"] = ((L.cout,), ("conv_b", L.init))
            if L.bn:
                out[f"{L.bn}.weight"] = ((L.cout,), ("bn_w", L.init))
                out[f"{L.bn}.bias"] = ((L.cout,), ("bn_b", L.init))
                out[f"{L.bn}.running_mean"] = ((L.cout,), ("bn_mean", L.init))
                out[f"{L.bn}.running_var"] = ((L.cout,), ("bn_var", L.init))
                out[f"{L.bn}.num_batches_tracked"] = ((), ("bn_count", L.init))
        return out


# ---------------------------------------------------------------------------
# Encoders (torchvision ResNet trunk as wrapped by smp ResNetEncoder [ext])
# ---------------------------------------------------------------------------
def _encoder(net: NetSpec, name: str, dilate_layer4: bool = False, dilations=None):
    """Returns the feature tensors [f1 (/2), f2 (/4), f3 (/8), f4 (/16), f5 (/32 or /16)].
    `dilations` {stage: rate} = smp `encoder.make_dilated(stage_list, dilation_list)` in ResNet stage
    numbering (layer3 = 3, layer4 = 4): every Conv2d of the stage gets stride 1, dilation `rate`,
    padding (k // 2) * rate."""
    dilations = dict(dilations or {})
    if dilate_layer4:
        dilations[4] = 2
    if name not in ENCODER_CFG:
        raise ValueError(f"encoder {name!r} is not supported by the B200 engine; options: {sorted(ENCODER_CFG)}")
    kind, blocks, groups, wpg = ENCODER_CFG[name]
    f1 = net.conv("encoder.conv1", [0], 64, 7, stride=2, pad=3, bn="encoder.bn1", relu=True)
    x = net.maxpool(f1)
    feats = [f1]
    inplanes = 64
    expansion = 1 if kind == "basic" else 4
    for stage, (planes, nblk) in enumerate(zip((64, 128, 256, 512), blocks), start=1):
        stage_stride = 1 if stage == 1 else 2
        dil = 1
        if stage in dilations:
            # smp replace_strides_with_dilation: every conv of the stage -> stride 1,
            # dilation r, padding (k // 2) * r
            stage_stride, dil = 1, dilations[stage]
        for b in range(nblk):
            p = f"encoder.layer{stage}.{b}"
            stride = stage_stride if b == 0 else 1
            need_ds = b == 0 and (stage != 1 or kind == "bottleneck")
            identity = x
            if need_ds:
                identity = net.conv(f"{p}.downsample.0", [x], planes * expansion, 1, stride=stride,
                                    bn=f"{p}.downsample.1")
            if kind == "basic":
                y = net.conv(f"{p}.conv1", [x], planes, 3, stride=stride, pad=dil, dil=dil,
                             bn=f"{p}.bn1", relu=True)
                x = net.conv(f"{p}.conv2", [y], planes, 3, pad=dil, dil=dil, bn=f"{p}.bn2",
                             relu=True, res=identity)
            else:
                width = int(planes * (wpg / 64.0)) * groups
                y = net.conv(f"{p}.conv1", [x], width, 1, bn=f"{p}.bn1", relu=True)
                y = net.conv(f"{p}.conv2", [y], width, 3, stride=stride, pad=dil, dil=dil,
                             groups=groups, bn=f"{p}.bn2", relu=True)
                x = net.conv(f"{p}.conv3", [y], planes * expansion, 1, bn=f"{p}.bn3", relu=True,
                             res=identity)
            inplanes = planes * expansion
        feats.append(x)
    return feats


def _decoder_block(net, prefix, x, skips, cout):
    """smp DecoderBlock [ext]: nearest x2 -> cat(skips) -> (conv3x3-BN-ReLU) x 2."""
    srcs = [(x, 1)] + [(s, 0) for s in skips]
    y = net.conv(f"{prefix}.conv1.0", srcs, cout, 3, pad=1, bn=f"{prefix}.conv1.1", relu=True,
                 init="decoder")
    return net.conv(f"{prefix}.conv2.0", [y], cout, 3, pad=1, bn=f"{prefix}.conv2.1", relu=True,
                    init="decoder")


def build_unet(encoder: str, classes: int, in_channels: int = 1) -> NetSpec:
    net = NetSpec("unet", encoder, classes, in_channels)
    f = _encoder(net, encoder)[::-1]  # f5, f4, f3, f2, f1
    x = f[0]
    for i, cout in enumerate(DECODER_CHANNELS):
        skips = [f[i + 1]] if i + 1 < len(f) else []
        x = _decoder_block(net, f"decoder.blocks.{i}", x, skips, cout)
    logits = net.conv("segmentation_head.0", [x], classes, 3, pad=1, bias=True, init="head",
                      out_dtype=1)
    net.head(logits)
    return net


def build_unetplusplus(encoder: str, classes: int, in_channels: int = 1) -> NetSpec:
    net = NetSpec("unetplusplus", encoder, classes, in_channels)
    f = _encoder(net, encoder)[::-1]
    ch = [net.tensors[t].channels for t in f]
    ins = [ch[0]] + list(DECODER_CHANNELS[:-1])
    skips_ch = ch[1:] + [0]
    outs = list(DECODER_CHANNELS)
    depth = len(ins) - 1
    cout_of = {}
    for l in range(len(ins) - 1):
        for d in range(l + 1):
            cout_of[(d, l)] = outs[l] if d == 0 else skips_ch[l]
    cout_of[(0, depth)] = outs[-1]
    dense = {}
    for l in range(len(ins) - 1):
        for d in range(depth - l):
            if l == 0:
                dense[(d, d)] = _decoder_block(net, f"decoder.blocks.x_{d}_{d}", f[d], [f[d + 1]],
                                               cout_of[(d, d)])
            else:
                li = d + l
                cat = [dense[(i, li)] for i in range(d + 1, li + 1)] + [f[li + 1]]
                dense[(d, li)] = _decoder_block(net, f"decoder.blocks.x_{d}_{li}",
                                                dense[(d, li - 1)], cat, cout_of[(d, li)])
    x = _decoder_block(net, f"decoder.blocks.x_0_{depth}", dense[(0, depth - 1)], [], cout_of[(0, depth)])
    logits = net.conv("segmentation_head.0", [x], classes, 3, pad=1, bias=True, init="head",
                      out_dtype=1)
    net.head(logits)
    return net


def build_deeplabv3plus(encoder: str, classes: int, in_channels: int = 1) -> NetSpec:
    net = NetSpec("deeplabv3plus", encoder, classes, in_channels)
    f = _encoder(net, encoder, dilate_layer4=True)  # f1..f5 ; f5 at /16
    top, hi = f[-1], f[-4]
    a = "decoder.aspp.0"
    branches = [net.conv(f"{a}.convs.0.0", [top], 256, 1, bn=f"{a}.convs.0.1", relu=True, init="decoder")]
    cin = net.tensors[top].channels
    for i, rate in enumerate((12, 24, 36), start=1):
        dw = net.conv(f"{a}.convs.{i}.0.0", [top], cin, 3, pad=rate, dil=rate, groups=cin, init="decoder")
        branches.append(net.conv(f"{a}.convs.{i}.0.1", [dw], 256, 1, bn=f"{a}.convs.{i}.1", relu=True,
                                 init="decoder"))
    pooled = net.gap(top)
    pc = net.conv(f"{a}.convs.4.1", [pooled], 256, 1, bn=f"{a}.convs.4.2", relu=True, init="decoder")
    branches.append(net.upsample(pc, net.tensors[top].ds_log2, mode=1))
    proj = net.conv(f"{a}.project.0", branches, 256, 1, bn=f"{a}.project.1", relu=True, init="decoder")
    dw = net.conv("decoder.aspp.1.0", [proj], 256, 3, pad=1, groups=256, init="decoder")
    asp = net.conv("decoder.aspp.1.1", [dw], 256, 1, bn="decoder.aspp.2", relu=True, init="decoder")
    up = net.upsample(asp, net.tensors[asp].ds_log2 - 2, mode=0, factor=4)
    hr = net.conv("decoder.block1.0", [hi], 48, 1, bn="decoder.block1.1", relu=True, init="decoder")
    cat_c = 256 + 48
    dw2 = net.conv("decoder.block2.0.0", [up, hr], cat_c, 3, pad=1, groups=cat_c, init="decoder")
    fused = net.conv("decoder.block2.0.1", [dw2], 256, 1, bn="decoder.block2.1", relu=True, init="decoder")
    logits = net.conv("segmentation_head.0", [fused], classes, 1, bias=True, init="head", out_dtype=1)
    net.head(logits, factor=4)
    return net


def build_deeplabv3(encoder: str, classes: int, in_channels: int = 1) -> NetSpec:
    """smp.DeepLabV3 [ext]: encoder at output stride 8 (make_dilated(stage_list=[4, 5], dilation_list=[2, 4]) in smp's
    stage numbering = layer3 dilation 2, layer4 dilation 4), DeepLabV3Decoder = Sequential(ASPP with plain 3x3 atrous
    convolutions at rates 12 / 24 / 36, conv3x3 + BN + ReLU), head = conv1x1 + bilinear x8 (align_corners=True)."""
    net = NetSpec("deeplabv3", encoder, classes, in_channels)
    f = _encoder(net, encoder, dilations={3: 2, 4: 4})
    top = f[-1]
    a = "decoder.0"
    branches = [net.conv(f"{a}.convs.0.0", [top], 256, 1, bn=f"{a}.convs.0.1", relu=True, init="decoder")]
    for i, rate in enumerate((12, 24, 36), start=1):
        branches.append(net.conv(f"{a}.convs.{i}.0", [top], 256, 3, pad=rate, dil=rate, bn=f"{a}.convs.{i}.1", relu=True,
                                 init="decoder"))
    pooled = net.gap(top)
    pc = net.conv(f"{a}.convs.4.1", [pooled], 256, 1, bn=f"{a}.convs.4.2", relu=True, init="decoder")
    branches.append(net.upsample(pc, net.tensors[top].ds_log2, mode=1))
    proj = net.conv(f"{a}.project.0", branches, 256, 1, bn=f"{a}.project.1", relu=True, init="decoder")
    y = net.conv("decoder.1", [proj], 256, 3, pad=1, bn="decoder.2", relu=True, init="decoder")
    logits = net.conv("segmentation_head.0", [y], classes, 1, bias=True, init="head", out_dtype=1)
    net.head(logits, factor=8)
    return net


BUILDERS = {
    "U_NET": build_unet,
    "DEEPLABV3": build_deeplabv3,
    "U_NET_PLUS_PLUS": build_unetplusplus,
    "DEEPLABV3_PLUS": build_deeplabv3plus,
}


def build_netspec(model_type_name: str, encoder_name: str, classes: int, in_channels: int = 1) -> NetSpec:
    if model_type_name not in BUILDERS:
        raise NotImplementedError(
            f"model type {model_type_name} is not supported by the B200 engine "
            f"(supported: {sorted(BUILDERS)}; SURVEY.md 8f-4 lists the rest as future work)"
        )
    return BUILDERS[model_type_name](encoder_name, classes, in_channels)
