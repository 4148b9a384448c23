"""Multi-GPU correctness in the driver-run suite: skipped below two GPUs.

* ``VolSeg2dPredictor`` with the additive ``cuda_devices`` setting (one process, N engines:
  sharded upload + NVLink all-gather of the volume, split work items, fused peer
  max-reduce + unpack, per-GPU download of its shard) equals the single-device prediction
  BIT FOR BIT -- labels and fp16 probabilities, 3-way and 12-way, ragged odd-sized volume.
* The one-process-per-GPU path bench.py uses under torchrun (CUDA-IPC peer exchange and the NCCL
  max all-reduce, tests/multi_gpu_check.py) reproduces the single-GPU result bit for bit; its log
  is kept under gpurun_out/ when that directory exists.
"""
import os
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
needs2 = pytest.mark.skipif(NGPU < 2, reason="needs at least two GPUs (gpurun --gpus 2)")

SETTINGS = dict(quality="high", output_probs=True, clip_data=False, st_dev_factor=2.575,
                data_hdf5_path="/data", cuda_device=0, downsample=False, one_hot=False, prediction_axis="Z")


@needs2
@pytest.mark.parametrize("shape", [(45, 96, 83), (64, 64, 64)])
def test_predictor_on_n_devices_equals_one_device(tmp_path, unet_r34, shape):
    import volume_segmantics.utilities.base_data_utils as utils
    from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor

    struc = {"type": utils.ModelType.U_NET, "encoder_name": "resnet34", "encoder_weights": None,
             "in_channels": 1, "classes": 4}
    path = tmp_path / "m.pytorch"
    torch.save({"model_state_dict": unet_r34[0].state_dict(), "model_struc_dict": struc, "label_codes": {}}, path)
    vol = np.random.default_rng(9).integers(0, 256, shape, dtype=np.uint8)
    one = VolSeg2dPredictor(str(path), SimpleNamespace(**SETTINGS))
    for n in sorted({2, NGPU}):
        many = VolSeg2dPredictor(str(path), SimpleNamespace(**dict(SETTINGS, cuda_devices=list(range(n)))))
        for fn in ("_predict_3_ways_max_probs", "_predict_12_ways_max_probs"):
            l1, p1 = getattr(one, fn)(vol)
            ln, pn = getattr(many, fn)(vol)
            assert ln.shape == shape and ln.dtype == np.uint8 and pn.dtype == np.float16
            assert np.array_equal(l1, ln), f"{n} GPUs {fn}: labels differ"
            assert np.array_equal(p1.view(np.uint16), pn.view(np.uint16)), f"{n} GPUs {fn}: probabilities differ"
        l1, none = one._predict_single_axis(vol, output_probs=False)
        ln, none_n = many._predict_single_axis(vol, output_probs=False)
        assert none is None and none_n is None and np.array_equal(l1, ln)


@needs2
def test_torchrun_sharded_exchange_is_bit_exact():
    n = 2 if NGPU < 4 else (4 if NGPU < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29631", str(ROOT / "tests" / "multi_gpu_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    log = res.stdout + "\n--- stderr ---\n" + res.stderr[-4000:]
    out_dir = ROOT / "gpurun_out"
    if out_dir.is_dir():
        (out_dir / f"r02_multi_gpu_check_{n}gpu.log").write_text(log)
    assert res.returncode == 0, log
    assert "MULTI-GPU CHECK PASSED" in res.stdout, log
