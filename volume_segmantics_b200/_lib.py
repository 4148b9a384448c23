"""ctypes binding of libvsb200.so (the C ABI in include/vsb200.h).

The library is built in-tree by ``volume_segmantics_b200.build``.  There is no
fallback: if the shared object is missing, cannot be loaded, or lacks a symbol
the header declares, importing the binding raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
# VSB200_VARIANT=bf16 selects the bfloat16 build; default is the fp16 build
# (same tcgen05 rate, 8x finer mantissa -- see DESIGN.md "numeric format").
VARIANT = os.environ.get("VSB200_VARIANT", "f16")
LIB_PATH = HERE / ("libvsb200.so" if VARIANT == "f16" else f"libvsb200_{VARIANT}.so")

VSB_MAX_SRC = 6
VSB_OP_CONV, VSB_OP_MAXPOOL, VSB_OP_GAP, VSB_OP_UPSAMPLE, VSB_OP_HEAD = 1, 2, 3, 4, 5

PROF_CLASSES = ("slicer", "conv_tc", "conv_simt", "stem", "pool", "head", "other")


class TensorDesc(C.Structure):
    _fields_ = [("channels", C.c_int32), ("ds_log2", C.c_int32), ("dtype", C.c_int32), ("reserved", C.c_int32)]


class Op(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("out", C.c_int32),
        ("n_src", C.c_int32),
        ("src", C.c_int32 * VSB_MAX_SRC),
        ("src_up", C.c_int32 * VSB_MAX_SRC),
        ("res", C.c_int32),
        ("cin", C.c_int32),
        ("cout", C.c_int32),
        ("kh", C.c_int32),
        ("kw", C.c_int32),
        ("stride", C.c_int32),
        ("pad", C.c_int32),
        ("dil", C.c_int32),
        ("groups", C.c_int32),
        ("relu", C.c_int32),
        ("mode", C.c_int32),
        ("factor", C.c_int32),
        ("w_off", C.c_int64),
        ("b_off", C.c_int64),
    ]


class Direction(C.Structure):
    _fields_ = [
        (n, C.c_int64)
        for n in (
            "S", "H", "W", "Hp", "Wp", "pad_top", "pad_left", "crop_top", "crop_left",
            "base", "stride_s", "stride_r", "stride_c",
        )
    ]


class VsbError(RuntimeError):
    pass


# name -> (restype, argtypes); must list every symbol include/vsb200.h declares
_P = C.c_void_p
SIGNATURES = {
    "vsb_abi_version": (C.c_int, []),
    "vsb_act_dtype": (C.c_int, []),
    "vsb_last_error": (C.c_char_p, []),
    "vsb_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "vsb_destroy": (None, [_P]),
    "vsb_load_plan": (C.c_int, [_P, C.POINTER(TensorDesc), C.c_int32, C.POINTER(Op), C.c_int32, _P, C.c_size_t, C.c_int32]),
    "vsb_direction_geometry": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.POINTER(Direction)]),
    "vsb_set_volume": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int64]),
    "vsb_set_volume_typed": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64]),
    "vsb_volume_generation": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "vsb_reset_keys": (C.c_int, [_P]),
    "vsb_predict_range": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64]),
    "vsb_predict": (C.c_int, [_P, C.c_uint32, C.c_int32]),
    "vsb_keys": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "vsb_bind_keys": (C.c_int, [_P, _P]),
    "vsb_keys_ipc_export": (C.c_int, [_P, _P]),
    "vsb_peers_open": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "vsb_peers_close": (C.c_int, [_P]),
    "vsb_peers_attach": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "vsb_set_volume_shard": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "vsb_volume_pull": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "vsb_fetch_shard": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "vsb_reduce_unpack_shard": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "vsb_fetch": (C.c_int, [_P, _P, _P]),
    "vsb_unpack_device": (C.c_int, [_P, _P, _P]),
    "vsb_set_vote_mode": (C.c_int, [_P, C.c_int32]),
    "vsb_fetch_votes": (C.c_int, [_P, _P]),
    "vsb_synchronize": (C.c_int, [_P]),
    "vsb_set_stream": (C.c_int, [_P, _P]),
    "vsb_launch_count": (C.c_int, [_P, C.POINTER(C.c_int64), C.c_int32]),
    "vsb_set_batch": (C.c_int, [_P, C.c_int32]),
    "vsb_set_conv_impl": (C.c_int, [_P, C.c_int32]),
    "vsb_set_flag": (C.c_int, [_P, C.c_char_p, C.c_int32]),
    "vsb_slice_batch": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int32, _P]),
    "vsb_slice_batch_generic": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int32, _P]),
    "vsb_merge_injected": (C.c_int, [_P, C.c_int32, _P, _P]),
    "vsb_forward_logits": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vsb_debug_tensor": (C.c_int, [_P, C.c_int32, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "vsb_stage_ms": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "vsb_clip_to_uint8": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_double, _P]),
    "vsb_raw_upload": (C.c_int, [_P, _P, C.c_int32, C.c_int64]),
    "vsb_raw_moments": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "vsb_raw_clip_to_volume": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int64, C.c_int64, _P,
                                         C.POINTER(C.c_uint64)]),
    "vsb_raw_release": (C.c_int, [_P]),
    "vsb_set_profiling": (C.c_int, [_P, C.c_int32]),
    "vsb_op_ms": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libvsb200.so and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VsbError(
            f"{LIB_PATH} not found: build it with `python -m volume_segmantics_b200.build` "
            "(the B200 engine has no CPU or PyTorch fallback)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is absent
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().vsb_last_error()
        raise VsbError(f"libvsb200 error {rc}: {msg.decode() if msg else ''}")


def act_dtype():
    """torch dtype of activations/weights in the loaded library build."""
    import torch

    return torch.float16 if load().vsb_act_dtype() == 1 else torch.bfloat16


def direction_geometry(Z: int, Y: int, X: int, d: int) -> Direction:
    g = Direction()
    check(load().vsb_direction_geometry(Z, Y, X, d, C.byref(g)))
    return g
