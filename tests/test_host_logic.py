"""Host-side logic that needs no GPU: enums / pickled model format, settings,
BatchNorm folding, plan lowering, work partitioning, key packing."""
import pickle
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle.smp_models import make_random_model
from volume_segmantics_b200 import _lib, sharding
from volume_segmantics_b200.plan import B200SegmentationModel, _fold, conv_macs_per_pixel, lower_to_plan


def test_reference_import_paths_and_enum_pickling():
    import volume_segmantics.utilities.base_data_utils as utils
    from volume_segmantics.data import get_settings_data
    from volume_segmantics.model import VolSeg2DPredictionManager
    from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor
    from volume_segmantics.utilities import Quality, get_2d_prediction_parser

    assert Quality.HIGH.value == 12 and utils.Axis.X.value == 2 and utils.ModelType.DEEPLABV3_PLUS.value == 5
    blob = pickle.dumps(utils.ModelType.U_NET)
    assert b"volume_segmantics.utilities.base_data_utils" in blob
    assert pickle.loads(blob) is utils.ModelType.U_NET
    assert callable(get_settings_data) and callable(get_2d_prediction_parser)
    assert VolSeg2DPredictionManager and VolSeg2dPredictor


def test_settings_loader(tmp_path):
    from volume_segmantics.data import get_settings_data

    p = tmp_path / "s.yaml"
    p.write_text("quality: high\ncuda_device: 0\none_hot: False\n")
    s = get_settings_data(p)
    assert s.quality == "high" and s.cuda_device == 0
    assert get_settings_data({"a": 1}).a == 1
    assert get_settings_data(None) == SimpleNamespace()
    with pytest.raises(SystemExit) as ex:  # reference tests/test_settings_data.py
        get_settings_data(tmp_path / "missing.yaml")
    assert ex.value.code == 1


def test_invalid_quality_exits_1():
    import volume_segmantics.utilities.base_data_utils as utils

    with pytest.raises(SystemExit) as ex:
        utils.get_prediction_quality(SimpleNamespace(quality="ultra"))
    assert ex.value.code == 1
    assert utils.get_prediction_axis(SimpleNamespace()) == utils.Axis.Z


def test_model_file_roundtrip(tmp_path):
    """The .pytorch dict written like reference tests/conftest.py:184-190 loads."""
    import volume_segmantics.utilities.base_data_utils as utils
    from volume_segmantics.model.model_2d import create_model_from_file

    oracle = make_random_model("unet", "resnet34", 4, seed=3)
    struc = {"type": utils.ModelType.U_NET, "encoder_name": "resnet34", "encoder_weights": "imagenet",
             "in_channels": 1, "classes": 4}
    path = tmp_path / "m.pytorch"
    torch.save({"model_state_dict": oracle.state_dict(), "model_struc_dict": struc, "label_codes": {"a": 1}}, path)
    model, n, codes = create_model_from_file(path)
    assert isinstance(model, torch.nn.Module) and n == 4 and codes == {"a": 1}
    for k, v in oracle.state_dict().items():
        assert torch.equal(model.state_dict()[k], v)


@pytest.mark.parametrize("mt,arch,enc,c,macs", [
    ("U_NET", "unet", "resnet34", 4, 118096),
    ("U_NET", "unet", "resnet34", 2, 117808),
    ("U_NET_PLUS_PLUS", "unetplusplus", "resnext50_32x4d", 6, 878448),
    ("DEEPLABV3_PLUS", "deeplabv3plus", "resnet50", 4, 137948),
    ("DEEPLABV3", "deeplabv3", "resnet34", 3, 415004),
    ("DEEPLABV3", "deeplabv3", "resnet50", 3, 622364),
])
def test_state_dict_keys_and_mac_counts(mt, arch, enc, c, macs):
    """smp key names == oracle key names; MAC/px equals SURVEY.md 8a-T."""
    from oracle.smp_models import OracleSegModel

    m = B200SegmentationModel(mt, enc, c)
    o = OracleSegModel(arch, enc, c)
    assert set(m.state_dict()) == set(o.state_dict())
    m.load_state_dict(o.state_dict())
    assert conv_macs_per_pixel(m.spec) == macs


def test_bn_folding_matches_torch():
    oracle = make_random_model("unet", "resnet34", 4, seed=0)
    m = B200SegmentationModel("U_NET", "resnet34", 4)
    m.load_state_dict(oracle.state_dict())
    L = next(l for l in m.spec.layers if l.name == "encoder.layer2.0.conv1")
    w_bits, b = _fold(m.state_dict(), L)
    w = torch.from_numpy(w_bits.view(np.int16).copy()).view(_lib.act_dtype()).float().permute(0, 3, 1, 2)
    x = torch.randn(1, 64, 9, 9)
    blk = oracle.encoder.layer2[0]
    with torch.no_grad():
        want = blk.bn1(blk.conv1(x))
        got = torch.nn.functional.conv2d(x, w, torch.from_numpy(b), stride=2, padding=1)
    assert torch.allclose(got, want, atol=3e-2, rtol=1e-2)  # 16-bit weight rounding only


def test_plan_lowering_tables():
    m = B200SegmentationModel("U_NET", "resnet34", 4)
    p = lower_to_plan(m)
    kinds = [op.kind for op in p.ops]
    assert kinds.count(_lib.VSB_OP_CONV) == 47 and kinds.count(_lib.VSB_OP_MAXPOOL) == 1
    assert kinds[-1] == _lib.VSB_OP_HEAD
    dec = [op for op in p.ops if op.kind == _lib.VSB_OP_CONV and op.n_src == 2]
    assert len(dec) == 4 and all(op.src_up[0] == 1 and op.src_up[1] == 0 for op in dec)
    for op in p.ops:
        if op.kind == _lib.VSB_OP_CONV:
            assert op.w_off % 256 == 0 and op.b_off % 256 == 0
            assert op.w_off + op.cout * op.kh * op.kw * (op.cin // op.groups) * 2 <= p.blob.size


def test_space_to_depth_lowering_of_upsample_concat_conv():
    """plan.s2d_up_concat_weights: cat(nearest-x2 upsample(x), skip) -> conv3x3  ==  conv3x3 at the resolution of
    x over [x, space-to-depth(skip)] followed by depth-to-space (the form conv_halo_el_kernel runs)."""
    import torch.nn.functional as F

    from volume_segmantics_b200.plan import s2d_up_concat_weights

    torch.manual_seed(0)
    o, cu, cs = 8, 6, 5
    w = torch.randn(o, cu + cs, 3, 3)
    x, skip = torch.randn(2, cu, 5, 7), torch.randn(2, cs, 10, 14)
    want = F.conv2d(torch.cat([F.interpolate(x, scale_factor=2, mode="nearest"), skip], 1), w, padding=1)

    def s2d(t):  # [N, C, 2H, 2W] -> [N, 4C, H, W], channel = (a*2 + b)*C + c for pixel (2i+a, 2j+b)
        n, c, h, wd = t.shape
        return t.reshape(n, c, h // 2, 2, wd // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, h // 2, wd // 2)

    got = F.conv2d(torch.cat([x, s2d(skip)], 1), s2d_up_concat_weights(w, cu), padding=1)
    n, _, h, wd = got.shape
    got = got.reshape(n, 2, 2, o, h, wd).permute(0, 3, 4, 1, 5, 2).reshape(n, o, 2 * h, 2 * wd)
    assert torch.allclose(got, want, atol=1e-5)


def test_plan_marks_decoder_conv1_for_the_space_to_depth_kernel():
    p = lower_to_plan(B200SegmentationModel("U_NET", "resnet34", 4))
    dec = [op for op in p.ops if op.kind == _lib.VSB_OP_CONV and op.n_src == 2]
    assert [op.mode for op in dec] == [2, 2, 2, 2]
    for op in dec:
        c_up = p.tensors[op.src[0]].channels
        k2 = c_up + 4 * (op.cin - c_up)
        assert op.factor > 0 and op.factor * 256 + 4 * op.cout * 9 * k2 * 2 <= p.blob.size
    assert all(op.mode == 0 for op in p.ops if op.kind == _lib.VSB_OP_CONV and op.n_src == 1)
    # U-Net++: every dense node with >= 32 output channels (several skip sources); DeepLabV3+: none (no up-sampled concat)
    pp = lower_to_plan(B200SegmentationModel("U_NET_PLUS_PLUS", "resnext50_32x4d", 6))
    marked = [op for op in pp.ops if op.kind == _lib.VSB_OP_CONV and op.mode == 2]
    assert len(marked) >= 9 and all(op.src_up[0] == 1 and not any(op.src_up[1:op.n_src]) and op.cout >= 32 for op in marked)
    assert max(op.n_src for op in marked) >= 4
    dl = lower_to_plan(B200SegmentationModel("DEEPLABV3_PLUS", "resnet50", 4))
    assert not [op for op in dl.ops if op.kind == _lib.VSB_OP_CONV and op.mode == 2]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("shape", [(64, 64, 64), (20, 40, 45), (2048, 2048, 512)])
def test_partition_covers_everything_once(world, shape):
    dirs = sharding.direction_list((1 << 12) - 1, skip_duplicates=True)
    assert dirs == [0, 1, 2, 4, 5, 7, 8, 11]
    shares = sharding.partition(shape, dirs, world)
    seen = {d: np.zeros(sharding.direction_dims(shape, d)[0], int) for d in dirs}
    for sh in shares:
        for it in sh:
            seen[it.d][it.s0:it.s1] += 1
    assert all((v == 1).all() for v in seen.values())
    costs = [sum(it.cost for it in sh) for sh in shares]
    assert max(costs) <= 1.15 * (sum(costs) / world) + max(it.cost for sh in shares for it in sh) / max(1, min(it.slices for sh in shares for it in sh)) * 8


def test_key_packing_reproduces_first_max_rule():
    rng = np.random.default_rng(0)
    n = 4096
    pal = np.array([0.9999, 0.99995, 1.0, 0.5, 0.50001, 0.25], np.float32)
    p = pal[rng.integers(0, len(pal), (12, n))]
    lab = rng.integers(0, 7, (12, n)).astype(np.uint8)
    keys = np.zeros(n, np.uint64)
    for d in rng.permutation(12):  # any order: max is commutative
        keys = np.maximum(keys, sharding.pack_keys_np(p[d], lab[d], int(d)))
    got_l, got_p = sharding.unpack_keys_np(keys)
    p16 = p.astype(np.float16)
    idx = np.argmax(p16, axis=0)  # numpy first-max, as _merge_vols_in_mem
    assert np.array_equal(got_l, lab[idx, np.arange(n)])
    assert np.array_equal(got_p, p16[idx, np.arange(n)])
    assert (keys < np.uint64(1) << np.uint64(63)).all()


def test_weights_version_tracks_in_place_updates_and_new_modules():
    """the plan key must change when weights are loaded / trained in place and must not
    depend on id() (CPython reuses ids after garbage collection)."""
    from volume_segmantics_b200.engine import weights_version

    a = B200SegmentationModel("U_NET", "resnet34", 2)
    v0 = weights_version(a)
    assert weights_version(a) == v0
    a.load_state_dict({k: t.clone() for k, t in a.state_dict().items()})
    v1 = weights_version(a)
    assert v1 != v0
    with torch.no_grad():
        next(a.parameters()).mul_(1.0)
    assert weights_version(a) != v1
    b = B200SegmentationModel("U_NET", "resnet34", 2)
    assert weights_version(b) != weights_version(a)


def test_in_channels_other_than_one_is_refused():
    with pytest.raises(NotImplementedError, match="single-channel"):
        B200SegmentationModel("U_NET", "resnet34", 2, in_channels=3)


def test_unsupported_volume_dtypes_are_named():
    from volume_segmantics_b200.host.predictor import _as_engine_volume

    assert _as_engine_volume(np.zeros((2, 4, 4), np.uint16)).dtype == np.uint16
    assert _as_engine_volume(np.zeros((2, 4, 4), np.int64)).dtype == np.int32
    assert _as_engine_volume(np.zeros((2, 4, 4), np.float32)).dtype == np.float32
    with pytest.raises(TypeError, match="float32"):
        _as_engine_volume(np.zeros((2, 4, 4), np.float64))
    with pytest.raises(ValueError, match="3-D"):
        _as_engine_volume(np.zeros((4, 4), np.uint8))
