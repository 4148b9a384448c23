"""Model container + lowering of a network to the libvsb200 op table.

``B200SegmentationModel`` is the ``nn.Module`` the drop-in ``VolSeg2dPredictor``
exposes as ``.model`` (reference vol_seg_2d_predictor.py:23-26 and
tests/test_vol_seg_2d_predictor.py:14-18 require an nn.Module).  It only *holds*
parameters under segmentation_models_pytorch's state-dict names; its forward
pass runs on the B200 engine (no torch compute, no CPU path).

``lower_to_plan`` folds every eval-mode BatchNorm into its convolution
(w' = w * g / sqrt(var + eps), b' = beta - mean * g / sqrt(var + eps)) in fp32,
rounds the weights to the library's 16-bit format in OHWI order and emits the ctypes tables.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .netspec import Layer, NetSpec, build_netspec

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, used by torchvision and smp


def _attach(root: nn.Module, dotted: str, value: torch.Tensor, is_param: bool) -> None:
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, nn.Module())
        mod = getattr(mod, p)
    if is_param:
        mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))
    else:
        mod.register_buffer(parts[-1], value)


class B200SegmentationModel(nn.Module):
    """Parameter container for one smp architecture (see netspec.py)."""

    def __init__(self, model_type_name: str, encoder_name: str, classes: int, in_channels: int = 1,
                 **_ignored):
        super().__init__()
        if int(in_channels) != 1:
            # the slicer writes ONE channel per voxel (vol_seg_2d_trainer.py:55 always builds in_channels=1);
            # a wider first convolution would read channels nobody wrote
            raise NotImplementedError(f"in_channels={in_channels}: the B200 engine slices single-channel volumes only")
        self.spec: NetSpec = build_netspec(model_type_name, encoder_name, classes, in_channels)
        self.model_type_name = model_type_name
        self.classes = classes
        self._engine = None  # set by the predictor; used by forward()
        for key, (shape, (role, init)) in self.spec.param_shapes().items():
            if role == "conv_w":
                w = torch.empty(shape)
                if init == "encoder":  # torchvision ResNet.__init__
                    nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
                elif init == "decoder":  # smp initialize_decoder
                    nn.init.kaiming_uniform_(w, mode="fan_in", nonlinearity="relu")
                else:  # smp initialize_head
                    nn.init.xavier_uniform_(w)
                _attach(self, key, w, True)
            elif role == "conv_b":
                _attach(self, key, torch.zeros(shape), True)
            elif role == "bn_w":
                _attach(self, key, torch.ones(shape), True)
            elif role == "bn_b":
                _attach(self, key, torch.zeros(shape), True)
            elif role == "bn_mean":
                _attach(self, key, torch.zeros(shape), False)
            elif role == "bn_var":
                _attach(self, key, torch.ones(shape), False)
            elif role == "bn_count":
                _attach(self, key, torch.tensor(0, dtype=torch.long), False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B,1,Hp,Wp] float (already padded + normalised) -> logits [B,C,Hp,Wp]
        computed by the B200 engine.  Raises without an engine: there is no
        torch implementation of the network in this package."""
        if self._engine is None:
            raise RuntimeError(
                "B200SegmentationModel.forward needs a bound B200 engine "
                "(construct it through VolSeg2dPredictor); no CPU/PyTorch fallback exists"
            )
        imgs = x.detach().to("cpu", torch.float32).numpy()[:, 0]
        logits = self._engine.forward_logits(self, np.ascontiguousarray(imgs))  # [B,Hp,Wp,C]
        return torch.from_numpy(logits).permute(0, 3, 1, 2).contiguous()


def _fold_f32(sd: Dict[str, torch.Tensor], L: Layer) -> Tuple[torch.Tensor, torch.Tensor]:
    """Conv weight [O,I,kh,kw] and bias with the eval-mode BatchNorm folded in (fp32)."""
    w = sd[f"{L.name}.weight"].detach().to("cpu", torch.float32)
    b = sd[f"{L.name}.bias"].detach().to("cpu", torch.float32) if L.has_bias else torch.zeros(L.cout)
    if L.bn:
        g = sd[f"{L.bn}.weight"].detach().to("cpu", torch.float32)
        beta = sd[f"{L.bn}.bias"].detach().to("cpu", torch.float32)
        mean = sd[f"{L.bn}.running_mean"].detach().to("cpu", torch.float32)
        var = sd[f"{L.bn}.running_var"].detach().to("cpu", torch.float32)
        scale = g / torch.sqrt(var + BN_EPS)
        w = w * scale[:, None, None, None]
        b = (b - mean) * scale + beta
    return w, b


def _pack(w: torch.Tensor, b: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 OIHW weights -> the library's 16-bit format in OHWI order; bias stays fp32."""
    w_ohwi = w.permute(0, 2, 3, 1).contiguous()
    if _lib.act_dtype() == torch.float16:
        w_ohwi = w_ohwi.clamp(-65504.0, 65504.0)
    w_ohwi = w_ohwi.to(_lib.act_dtype())
    return w_ohwi.view(torch.int16).numpy().view(np.uint16), b.numpy().astype(np.float32)


def _fold(sd: Dict[str, torch.Tensor], L: Layer) -> Tuple[np.ndarray, np.ndarray]:
    return _pack(*_fold_f32(sd, L))


# ---------------------------------------------------------------------------
# Space-to-depth ("S2D") tail.
#
# The last decoder block of smp's Unet / UnetPlusPlus is  nearest-x2 upsample ->
# conv3x3(Cin->16)+BN+ReLU -> conv3x3(16->16)+BN+ReLU, followed by the 3x3 segmentation
# head (16->classes), all at full resolution with GEMM N = 16: the tensor pipe idles
# while the shared-memory A operand streams.  The three convolutions are rewritten, with
# identical arithmetic up to summation order, as 3x3 convolutions at HALF resolution whose
# channels carry the 2x2 sub-pixels (channel = (2*a + b) * C + c for output pixel
# (2i+a, 2j+b)), so N becomes 4*16 = 64:
#   * upsample -> conv: every output sub-pixel sees the 3x3 low-resolution
#     neighbourhood; taps that land on the same source pixel are summed (in fp32, before
#     the one rounding to 16 bit) -- the same multiply-accumulate count as the original;
#   * conv on an S2D tensor: the 4x4 input patch of a 2x2 output block, spread over
#     the 3x3 S2D neighbourhood; 20 of the 36 (tap, sub-pixel) K-blocks are zero and the
#     engine skips them (it scans the packed weights for all-zero K-steps).
# The logits tensor is [Hp/2, Wp/2, 4*classes]; the HEAD op carries mode = 1 and the
# head kernel reads four voxels per thread.
# ---------------------------------------------------------------------------
def s2d_upconv_weights(w: torch.Tensor) -> torch.Tensor:
    """[O,I,3,3] of `nearest-x2 upsample -> conv3x3 pad 1`  ->  [4*O,I,3,3] conv3x3 pad 1 on the
    low-resolution source producing the S2D output."""
    o, i = w.shape[:2]
    out = torch.zeros(4, o, i, 3, 3, dtype=torch.float32)
    for a in (0, 1):
        for b in (0, 1):
            for ky in range(3):
                dy = (a + ky - 1) // 2  # source row offset of up-sampled row 2i+a+ky-1
                for kx in range(3):
                    dx = (b + kx - 1) // 2
                    out[a * 2 + b, :, :, dy + 1, dx + 1] += w[:, :, ky, kx]
    return out.reshape(4 * o, i, 3, 3)


def s2d_conv_weights(w: torch.Tensor) -> torch.Tensor:
    """[O,I,3,3] conv3x3 pad 1 at full resolution -> [4*O,4*I,3,3] conv3x3 pad 1 between S2D tensors."""
    o, i = w.shape[:2]
    out = torch.zeros(4, o, 4, i, 3, 3, dtype=torch.float32)
    for a in (0, 1):
        for b in (0, 1):
            for ky in range(3):
                r = a + ky - 1
                dy, a2 = r // 2, r % 2
                for kx in range(3):
                    c = b + kx - 1
                    dx, b2 = c // 2, c % 2
                    out[a * 2 + b, :, a2 * 2 + b2, :, dy + 1, dx + 1] = w[:, :, ky, kx]
    return out.reshape(4 * o, 4 * i, 3, 3)


def s2d_up_concat_weights(w: torch.Tensor, c_up: int) -> torch.Tensor:
    """[O, c_up + c_skip, 3, 3] of `cat(nearest-x2 upsample(x), skip) -> conv3x3 pad 1` (smp DecoderBlock.conv1)
    -> [4*O, c_up + 4*c_skip, 3, 3]: the same layer as a 3x3 convolution at the resolution of x, whose inputs are
    x itself and the space-to-depth view of the skip tensors and whose output is the space-to-depth view of the
    layer's output (vsb200.h, vsb_op.mode == 2)."""
    return torch.cat([s2d_upconv_weights(w[:, :c_up]), s2d_conv_weights(w[:, c_up:])], dim=1)


def s2d_up_layer(spec: NetSpec, L: Layer) -> bool:
    """Layers the engine can run in space-to-depth form (engine.cu prepare_el_plan)."""
    import os

    if os.environ.get("VSB200_S2D_UP", "1") == "0":
        return False
    if not (L.kind == "conv" and L.k == 3 and L.pad == 1 and L.stride == 1 and L.dil == 1 and L.groups == 1
            and L.res < 0 and len(L.srcs) >= 2 and L.srcs[0][1] and not any(up for _, up in L.srcs[1:])):
        return False
    if L.cout not in (32, 64, 128, 256) or spec.tensors[L.out].dtype or spec.tensors[L.out].ds_log2 < 0:
        return False
    if spec.tensors[L.srcs[0][0]].channels % 64:
        return False
    return all(spec.tensors[t].channels % 32 == 0 for t, _ in L.srcs[1:])


def find_s2d_tail(spec: NetSpec):
    """Indices (A, B, H, head) of the layers the S2D rewrite applies to, or None."""
    import os

    if os.environ.get("VSB200_S2D_TAIL", "1") == "0":
        return None
    layers = spec.layers
    if not layers or layers[-1].kind != "head" or layers[-1].factor != 1 or spec.classes > 8:
        return None
    producer = {L.out: i for i, L in enumerate(layers) if L.out >= 0}
    uses: Dict[int, int] = {}
    for L in layers:
        for t, _ in L.srcs:
            uses[t] = uses.get(t, 0) + 1
        if L.res >= 0:
            uses[L.res] = uses.get(L.res, 0) + 1

    def plain3x3(L: Layer) -> bool:
        return (L.kind == "conv" and L.k == 3 and L.pad == 1 and L.stride == 1 and L.dil == 1 and L.groups == 1
                and L.res < 0 and len(L.srcs) == 1)

    hi = producer.get(layers[-1].srcs[0][0])
    if hi is None or not plain3x3(layers[hi]) or layers[hi].srcs[0][1] or uses.get(layers[hi].out) != 1:
        return None
    bi = producer.get(layers[hi].srcs[0][0])
    if bi is None or not plain3x3(layers[bi]) or layers[bi].srcs[0][1] or uses.get(layers[bi].out) != 1:
        return None
    ai = producer.get(layers[bi].srcs[0][0])
    if ai is None or not plain3x3(layers[ai]) or not layers[ai].srcs[0][1] or uses.get(layers[ai].out) != 1:
        return None
    A, B = layers[ai], layers[bi]
    if A.cin % 16 or A.cout % 16 or B.cout % 16 or spec.tensors[A.out].ds_log2 != 0:
        return None
    if spec.tensors[A.out].dtype or spec.tensors[B.out].dtype or not spec.tensors[layers[hi].out].dtype:
        return None
    return ai, bi, hi, len(layers) - 1


class Plan:
    """ctypes tables handed to vsb_load_plan (kept alive by this object)."""

    def __init__(self, tensors, ops, blob: np.ndarray, classes: int):
        self.tensors, self.ops, self.blob, self.classes = tensors, ops, blob, classes


def lower_to_plan(model: B200SegmentationModel) -> Plan:
    spec = model.spec
    sd = model.state_dict()
    tail = find_s2d_tail(spec)
    tensors = (_lib.TensorDesc * len(spec.tensors))()
    for i, t in enumerate(spec.tensors):
        tensors[i] = _lib.TensorDesc(t.channels, t.ds_log2, t.dtype, 0)
    if tail:
        for li in tail[:3]:  # S2D tensors: 4x the channels at half the resolution
            t = spec.tensors[spec.layers[li].out]
            tensors[spec.layers[li].out] = _lib.TensorDesc(4 * t.channels, t.ds_log2 + 1, t.dtype, 0)
    ops = (_lib.Op * len(spec.layers))()
    chunks, off = [], 0

    def add(arr: np.ndarray) -> int:
        nonlocal off
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        pad = (-raw.size) % 256
        chunks.append(raw)
        if pad:
            chunks.append(np.zeros(pad, np.uint8))
        start = off
        off += raw.size + pad
        return start

    kinds = {"conv": _lib.VSB_OP_CONV, "maxpool": _lib.VSB_OP_MAXPOOL, "gap": _lib.VSB_OP_GAP,
             "upsample": _lib.VSB_OP_UPSAMPLE, "head": _lib.VSB_OP_HEAD}
    for i, L in enumerate(spec.layers):
        op = _lib.Op()
        op.kind = kinds[L.kind]
        op.out = L.out
        op.n_src = len(L.srcs)
        if op.n_src > _lib.VSB_MAX_SRC:
            raise ValueError(f"layer {L.name}: {op.n_src} sources > {_lib.VSB_MAX_SRC}")
        for j, (t, up) in enumerate(L.srcs):
            op.src[j], op.src_up[j] = t, up
        op.res = L.res
        op.w_off = op.b_off = -1
        op.mode, op.factor = L.mode, L.factor
        if L.kind == "conv":
            wf, bf = _fold_f32(sd, L)
            cin, cout = L.cin, L.cout
            if tail and i == tail[0]:
                wf, bf = s2d_upconv_weights(wf), bf.repeat(4)
                cout = 4 * cout
                op.src_up[0] = 0
            elif tail and i in tail[1:3]:
                wf, bf = s2d_conv_weights(wf), bf.repeat(4)
                cin, cout = 4 * cin, 4 * cout
            w, b = _pack(wf, bf)
            op.cin, op.cout, op.kh, op.kw = cin, cout, L.k, L.k
            op.stride, op.pad, op.dil, op.groups, op.relu = L.stride, L.pad, L.dil, L.groups, int(L.relu)
            op.w_off = add(w)
            op.b_off = add(b)
            if not (tail and i in tail[:3]) and s2d_up_layer(spec, L):
                w2, _ = _pack(s2d_up_concat_weights(wf, spec.tensors[L.srcs[0][0]].channels), bf)
                off2 = add(w2)
                assert off2 % 256 == 0
                op.mode, op.factor = 2, off2 // 256
        elif tail and i == tail[3]:
            op.mode = 1  # logits are S2D: [Hp/2, Wp/2, 4*classes]
        ops[i] = op
    blob = np.concatenate(chunks) if chunks else np.zeros(1, np.uint8)
    return Plan(tensors, ops, blob, spec.classes)


def conv_macs_per_pixel(spec: NetSpec) -> float:
    """Algorithmic multiply-accumulates per padded input pixel."""
    total = 0.0
    for L in spec.layers:
        if L.kind != "conv":
            continue
        ds = spec.tensors[L.out].ds_log2
        if ds < 0:
            continue
        total += L.k * L.k * (L.cin // L.groups) * L.cout / float(4 ** ds)
    return total
