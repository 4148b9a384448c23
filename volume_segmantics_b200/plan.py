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


def _fold(sd: Dict[str, torch.Tensor], L: Layer) -> Tuple[np.ndarray, np.ndarray]:
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
    w_ohwi = w.permute(0, 2, 3, 1).contiguous()
    if _lib.act_dtype() == torch.float16:
        w_ohwi = w_ohwi.clamp(-65504.0, 65504.0)
    w_ohwi = w_ohwi.to(_lib.act_dtype())
    return w_ohwi.view(torch.int16).numpy().view(np.uint16), b.numpy().astype(np.float32)


class Plan:
    """ctypes tables handed to vsb_load_plan (kept alive by this object)."""

    def __init__(self, tensors, ops, blob: np.ndarray, classes: int):
        self.tensors, self.ops, self.blob, self.classes = tensors, ops, blob, classes


def lower_to_plan(model: B200SegmentationModel) -> Plan:
    spec = model.spec
    sd = model.state_dict()
    tensors = (_lib.TensorDesc * len(spec.tensors))()
    for i, t in enumerate(spec.tensors):
        tensors[i] = _lib.TensorDesc(t.channels, t.ds_log2, t.dtype, 0)
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
            w, b = _fold(sd, L)
            op.cin, op.cout, op.kh, op.kw = L.cin, L.cout, L.k, L.k
            op.stride, op.pad, op.dil, op.groups, op.relu = L.stride, L.pad, L.dil, L.groups, int(L.relu)
            op.w_off = add(w)
            op.b_off = add(b)
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
