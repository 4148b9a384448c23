"""Both library builds are exercised by the driver-run suite: the default fp16
build by every other GPU test, the bfloat16 build (``VSB200_VARIANT=bf16``, the format north_star
names) by tests/variant_check.py in a sub-process -- the variant is bound when the library loads."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("variant", ["bf16", "f16"])
def test_library_variant(variant):
    env = dict(os.environ, VSB200_VARIANT=variant)
    res = subprocess.run([sys.executable, str(ROOT / "tests" / "variant_check.py")], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr[-3000:]
    assert f"VARIANT {variant} OK" in res.stdout
