"""Equivalence of the engine's alternative code paths on the GPU (each switch selects a
different kernel or pipeline for the same arithmetic):

* shared-memory + TMA-store epilogue, two MMA warps, two epilogue groups  vs  the direct epilogue /
  single-warp pipeline: same fp32 accumulation, bias, residual, ReLU and rounding -> bit-identical logits;
* fused stem + max-pool  vs  the separate max-pool kernel: max of the same fp16 values -> bit-identical;
* space-to-depth tail (plan.py)  vs  the plain last decoder block: merged weights are rounded once
  instead of per tap, so logits agree to 16-bit noise and labels almost everywhere -- checked for
  2, 3, 4 and 6 classes (the S2D head kernels are instantiated per class count), on a ragged shape and
  through all three axes (row and x-plane head kernels).
"""
import numpy as np
import pytest

from oracle import make_golden as mg
from oracle import predict_oracle as po
from oracle.smp_models import make_random_model
from volume_segmantics_b200.plan import B200SegmentationModel

pytestmark = pytest.mark.gpu


def _inputs(shape=(2, 70, 100), seed=11):
    vol = mg.structured_volume(shape, seed)
    return np.stack([po.preprocess_slice(vol[i]) for i in range(shape[0])]).astype(np.float32)


@pytest.mark.parametrize("flag,value", [("tma_epilogue", 0), ("mma_warps", 1), ("epi_groups", 0), ("fuse_pool", 0),
                                        ("halo_a_stages", 2), ("halo2_mma2", 1), ("halo_mt", 1), ("halo2_tma", 0), ("stem", 2), ("stem", 1), ("tc_smem_epilogue", 0),
                                        ("res_inplace", 0), ("el_tma_epilogue", 0)])
def test_pipeline_switches_are_bit_identical(engine, unet_r34, flag, value):
    _, model = unet_r34
    x = _inputs()
    want = engine.forward_logits(model, x)
    engine.set_flag(flag, value)
    try:
        got = engine.forward_logits(model, x)
    finally:
        engine.set_flag(flag, {"tma_epilogue": 1, "mma_warps": 2, "epi_groups": 1, "fuse_pool": 1,
                               "halo_a_stages": 8, "halo2_mma2": 0, "halo_mt": 2, "halo2_tma": 1, "stem": 3, "tc_smem_epilogue": 1,
                               "res_inplace": 1, "el_tma_epilogue": 1}[flag])
    assert np.array_equal(got, want), f"{flag}={value}: max |diff| {np.abs(got - want).max()}"


def test_pipeline_switches_bit_identical_on_larger_images(engine, unet_r34):
    """Many tiles per CTA (the ring / accumulator-stage bookkeeping wraps several times)."""
    _, model = unet_r34
    x = _inputs((3, 300, 420), 5)
    want = engine.forward_logits(model, x)
    for flag, value, back in (("tma_epilogue", 0, 1), ("mma_warps", 1, 2), ("halo_mt", 1, 2), ("stem", 2, 3), ("stem", 1, 3),
                              ("res_inplace", 0, 1), ("el_tma_epilogue", 0, 1)):
        engine.set_flag(flag, value)
        try:
            got = engine.forward_logits(model, x)
        finally:
            engine.set_flag(flag, back)
        assert np.array_equal(got, want), flag


@pytest.mark.parametrize("classes", [2, 3, 4, 6])
def test_s2d_tail_matches_plain_tail(engine, monkeypatch, classes):
    oracle = make_random_model("unet", "resnet34", classes, seed=3)
    vol = mg.structured_volume((9, 45, 70), 21)  # ragged: pad 19 / 26, crop offsets differ from pad offsets

    def run(s2d):
        monkeypatch.setenv("VSB200_S2D_TAIL", "1" if s2d else "0")
        model = B200SegmentationModel("U_NET", "resnet34", classes)
        model.load_state_dict(oracle.state_dict())
        engine.load_model(model)
        logits = engine.forward_logits(model, _inputs((2, 64, 96), 4))
        engine.set_volume(vol)
        engine.predict(0b111, True)
        labels, probs = engine.fetch()
        return logits, labels.copy(), probs.astype(np.float32)

    lg1, lab1, pr1 = run(True)
    lg0, lab0, pr0 = run(False)
    assert lg1.shape == lg0.shape
    assert np.abs(lg1 - lg0).max() < 5e-3 * max(1.0, np.abs(lg0).max())
    assert np.abs(pr1 - pr0).max() < 5e-3
    assert (lab1 == lab0).mean() > 0.995  # random-init weights: every voxel sits near a decision boundary


@pytest.mark.parametrize("mt,arch,enc,classes,shape", [
    ("U_NET", "unet", "resnet34", 4, (2, 150, 200)),       # several tiles per launch, overhanging tiles
    ("U_NET", "unet", "resnet34", 4, (3, 300, 420)),
    ("U_NET", "unet", "resnet50", 2, (2, 70, 100)),        # skip tensors of 256 / 512 / 1024 channels
    ("U_NET_PLUS_PLUS", "unetplusplus", "resnext50_32x4d", 6, (2, 70, 100)),  # several skip sources per layer
])
@pytest.mark.parametrize("flag", ["s2d_up", "el_conv"])
def test_s2d_up_concat_matches_parity_split(engine, mt, arch, enc, classes, shape, flag):
    """Decoder conv1 layers as space-to-depth convolutions (conv_halo_el_kernel, vsb_op.mode == 2)  vs  the
    parity-split kernels on the same plan: the up-sampled taps are summed before the one rounding to 16 bit
    and the products are accumulated in another order, so logits agree to 16-bit noise.
    "el_conv": the same kernel for 32 -> 32 convolutions (space-to-depth form) and stride-2 3x3 convolutions (parity
    planes of the input): identical products, another summation order."""
    oracle = make_random_model(arch, enc, classes, seed=6)
    model = B200SegmentationModel(mt, enc, classes)
    model.load_state_dict(oracle.state_dict())
    x = _inputs(shape, 13)
    got = engine.forward_logits(model, x)
    engine.set_flag(flag, 0)
    try:
        want = engine.forward_logits(model, x)
    finally:
        engine.set_flag(flag, 1)
    err = np.abs(got - want).max()
    print(f"[{flag} {mt}/{enc} {shape}] max |logit diff| {err:.2e} (|logits| max {np.abs(want).max():.3f}), "
          f"identical {np.array_equal(got, want)}")
    assert not np.array_equal(got, want), "the space-to-depth path did not run"
    assert err < 5e-3 * max(1.0, np.abs(want).max())


def test_s2d_head_against_oracle_three_axes(engine, unet_r34):
    """Z / Y axes use the row head kernel, X the x-plane kernel; probabilities within the BASELINE tolerance."""
    oracle, model = unet_r34
    engine.load_model(model)
    vol = mg.structured_volume((34, 40, 45), 8)  # > 32 slices along every axis: full and partial x-plane tiles
    engine.set_volume(vol)
    engine.predict(0b111, True)
    labels, probs = engine.fetch()
    ora = po.OraclePredictor(oracle, 4)
    want_l, want_p = ora.predict_3_ways_max_probs(vol)
    assert np.abs(probs.astype(np.float32) - want_p.astype(np.float32)).max() < 2e-2
    bad = labels != want_l
    cb = np.sort(ora.class_best_over_directions(vol, range(3)), axis=0)
    assert not bad.any() or (cb[-1] - cb[-2])[bad].max() < 2e-2  # margin clause
    print(f"[s2d head vs oracle] agreement {1 - bad.mean():.5f}")
    assert 1 - bad.mean() > 0.997  # random-init weights (near-ties everywhere): see tests/test_predictor_gpu.py


@pytest.mark.parametrize("mt,arch,enc,classes", [("U_NET", "unet", "resnet50", 2), ("DEEPLABV3_PLUS", "deeplabv3plus", "resnet50", 4)])
def test_per_tap_smem_epilogue_bit_identical_on_bottleneck_encoders(engine, mt, arch, enc, classes):
    """Wide 1x1 convolutions with a residual (bottleneck conv3, downsample): TMA-store epilogue == direct epilogue."""
    oracle = make_random_model(arch, enc, classes, seed=2)
    model = B200SegmentationModel(mt, enc, classes)
    model.load_state_dict(oracle.state_dict())
    x = _inputs((3, 150, 200), 9)
    want = engine.forward_logits(model, x)
    engine.set_flag("tc_smem_epilogue", 0)
    try:
        got = engine.forward_logits(model, x)
    finally:
        engine.set_flag("tc_smem_epilogue", 1)
    assert np.array_equal(got, want), f"max |diff| {np.abs(got - want).max()}"


@pytest.mark.parametrize("shape", [(2, 150, 200), (2, 520, 650)])
def test_depthwise_on_tensor_cores_matches_cuda_core_kernel(engine, shape):
    """DeepLabV3+ depthwise 3x3 convolutions (304 channels at 1/4 resolution: a last block of 48 channels; 256 and
    2048 channels at 1/16 with dilation 1 / 12 / 24 / 36) as block-diagonal tensor-core convolutions  vs  the
    CUDA-core depthwise kernel: same 16-bit weights and fp32 accumulation, exact zeros added, another order."""
    oracle = make_random_model("deeplabv3plus", "resnet50", 4, seed=4)
    model = B200SegmentationModel("DEEPLABV3_PLUS", "resnet50", 4)
    model.load_state_dict(oracle.state_dict())
    x = _inputs(shape, 17)
    got = engine.forward_logits(model, x)
    engine.set_flag("dw_tc", 0)
    try:
        want = engine.forward_logits(model, x)
    finally:
        engine.set_flag("dw_tc", 1)
    err = np.abs(got - want).max()
    print(f"[dw_tc {shape}] max |logit diff| {err:.2e} (|logits| max {np.abs(want).max():.3f}), identical {np.array_equal(got, want)}")
    assert err < 2e-3 * max(1.0, np.abs(want).max())
