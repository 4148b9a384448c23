"""Merge kernel (c'): injected per-direction probabilities/labels -> fp16 cast ->
inverse rotation -> packed-key atomicMax -> unpack, bit-exact against the numpy
restatement of vol_seg_2d_predictor.py:58-64,90-98,110-111 (engineered fp16 ties,
earliest direction wins)."""
import numpy as np
import pytest

from oracle import predict_oracle as po
from volume_segmantics_b200.sharding import direction_dims

pytestmark = pytest.mark.gpu

CASES = {"low": [0], "lowY": [1], "lowX": [2], "medium": [0, 1, 2], "high": list(range(12)),
         "high_nodup": [0, 1, 2, 4, 5, 7, 8, 11]}


@pytest.mark.parametrize("case", list(CASES))
def test_injected_merge_golden(engine, golden_dir, case):
    z = np.load(golden_dir / "merge_injected.npz")
    shape = z["high_labels"].shape
    engine.set_volume(np.zeros(shape, np.uint8))
    order = CASES[case][::-1]  # any order: the key carries the direction
    for d in order:
        engine.merge_injected(d, z[f"in_probs_{d}"], z[f"in_labels_{d}"])
    lab, prb = engine.fetch()
    assert np.array_equal(lab, z[f"{case}_labels"])
    assert np.array_equal(prb.view(np.uint16), z[f"{case}_probs"])


@pytest.mark.parametrize("shape", [(33, 47, 52), (64, 64, 64), (1, 5, 3)])
def test_injected_merge_random_with_ties(engine, shape):
    rng = np.random.default_rng(11)
    engine.set_volume(np.zeros(shape, np.uint8))
    pal = np.concatenate([np.linspace(0.2, 1.0, 40, dtype=np.float32), np.float32([0.9999, 0.99995, 0.33333, 0.33334])])
    probs = {d: pal[rng.integers(0, len(pal), direction_dims(shape, d))] for d in range(12)}
    labels = {d: rng.integers(0, 256, direction_dims(shape, d)).astype(np.uint8) for d in range(12)}
    for d in rng.permutation(12):
        engine.merge_injected(int(d), probs[int(d)], labels[int(d)])
    lab, prb = engine.fetch()
    want_l, want_p = po.merge_injected_oracle(shape, list(range(12)), probs, labels)
    assert np.array_equal(lab, want_l)
    assert np.array_equal(prb.view(np.uint16), want_p.view(np.uint16))


def test_merge_idempotent_and_reset(engine):
    shape = (16, 24, 40)
    rng = np.random.default_rng(3)
    engine.set_volume(np.zeros(shape, np.uint8))
    p = rng.random(direction_dims(shape, 5)).astype(np.float32)
    l = rng.integers(0, 9, direction_dims(shape, 5)).astype(np.uint8)
    engine.merge_injected(5, p, l)
    a = engine.fetch()
    engine.merge_injected(5, p, l)  # same direction again: max is idempotent
    b = engine.fetch()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    engine.reset()
    lab, prb = engine.fetch()
    assert not lab.any() and not prb.view(np.uint16).any()


def test_reduce_unpack_shard_single_rank_equals_fetch(engine):
    """The fused reduce+unpack kernel with one rank (no peers) is the plain unpack."""
    import torch

    shape = (9, 21, 35)  # odd voxel count: exercises the scalar tail
    rng = np.random.default_rng(4)
    engine.set_volume(np.zeros(shape, np.uint8))
    for d in (0, 5, 7):
        engine.merge_injected(d, rng.random(direction_dims(shape, d)).astype(np.float32),
                              rng.integers(0, 200, direction_dims(shape, d)).astype(np.uint8))
    want_l, want_p = engine.fetch()
    n = int(np.prod(shape))
    lab = torch.zeros(n, dtype=torch.uint8, device="cuda:0")
    prb = torch.zeros(n, dtype=torch.float16, device="cuda:0")
    engine.reduce_unpack_shard(0, n, lab.data_ptr(), prb.data_ptr())
    engine.synchronize()
    assert np.array_equal(lab.cpu().numpy(), want_l.ravel())
    assert np.array_equal(prb.cpu().numpy().view(np.uint16), want_p.ravel().view(np.uint16))
    # a sub-shard
    engine.reduce_unpack_shard(8, 1000, lab.data_ptr(), prb.data_ptr())
    engine.synchronize()
    assert np.array_equal(lab.cpu().numpy()[:992], want_l.ravel()[8:1000])
