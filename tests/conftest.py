import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def engine():
    """One libvsb200 engine on cuda:0 for the whole GPU session."""
    from volume_segmantics_b200.engine import Engine

    eng = Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def unet_r34():
    """(oracle model, B200 model) with identical seeded weights, BN stats randomised."""
    from oracle.smp_models import make_random_model
    from volume_segmantics_b200.plan import B200SegmentationModel

    oracle = make_random_model("unet", "resnet34", 4, seed=0)
    model = B200SegmentationModel("U_NET", "resnet34", 4)
    model.load_state_dict(oracle.state_dict())
    return oracle, model


@pytest.fixture(scope="session")
def trained_unet_r34():
    """(oracle model, B200 model) after ~40 Adam steps on synthetic labels, so the
    network is decisive like a real checkpoint (oracle/train_synth.py)."""
    import torch

    from oracle.train_synth import make_trained_model
    from volume_segmantics_b200.plan import B200SegmentationModel

    torch.set_num_threads(max(1, torch.get_num_threads()))
    oracle, _ = make_trained_model("unet", "resnet34", 4, seed=0, steps=40)
    model = B200SegmentationModel("U_NET", "resnet34", 4)
    model.load_state_dict(oracle.state_dict())
    return oracle, model


def act_bits(a, dtype=None):
    """uint16 bits of `a` rounded (RNE) to the library build's 16-bit format."""
    import torch

    from volume_segmantics_b200 import _lib

    dtype = dtype or _lib.act_dtype()
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dtype).view(torch.int16).numpy().view(np.uint16)


def act_tag():
    import torch

    from volume_segmantics_b200 import _lib

    return "f16" if _lib.act_dtype() == torch.float16 else "bf16"
