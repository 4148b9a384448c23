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


def bf16_bits(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
