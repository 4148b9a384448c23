"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/vsb200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "vsb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from volume_segmantics_b200 import _lib

    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vsb200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert lib.vsb_abi_version() == 2


def test_struct_layouts_match_header():
    from volume_segmantics_b200 import _lib

    assert ctypes.sizeof(_lib.TensorDesc) == 16
    assert ctypes.sizeof(_lib.Op) == 4 * (3 + 6 + 6 + 1 + 2 + 6 + 1 + 2) + 4 + 16  # +4 pad before int64
    assert ctypes.sizeof(_lib.Direction) == 13 * 8


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from volume_segmantics_b200 import _lib
    from volume_segmantics_b200.engine import Engine

    with pytest.raises(_lib.VsbError, match="no CPU fallback"):
        Engine(0)


def test_model_forward_has_no_torch_fallback():
    import torch
    from volume_segmantics_b200.plan import B200SegmentationModel

    m = B200SegmentationModel("U_NET", "resnet34", 2)
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        m(torch.zeros(1, 1, 32, 32))
