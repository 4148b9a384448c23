"""Generate the committed golden vectors under tests/golden/ from the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Run here (CPU container):  python -m oracle.make_golden
The reference itself cannot be imported in this image, so these
vectors pin the *oracle* (the restatement built on the same cv2 / torchvision /
numpy calls the reference reaches), and the GPU tests compare libvsb200 to them.
"""
from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np
import torch

from . import predict_oracle as po
from .smp_models import make_random_model

OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"


def act_bits(a: np.ndarray, dtype=torch.bfloat16) -> np.ndarray:
    """16-bit RNE encoding (bfloat16 or float16) of an fp32 array, as uint16 bits."""
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dtype).view(torch.int16).numpy().view(np.uint16)


def synth_volume(shape, seed):
    return np.random.default_rng(seed).integers(0, 256, size=shape, dtype=np.uint8)


def structured_volume(shape, seed):
    """three low-frequency sinusoids + N(0,20) noise, clipped to uint8."""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    v = 128 + 40 * np.sin(z / 5.0 + 0.3) + 35 * np.sin(y / 7.0 + 1.1) + 30 * np.sin(x / 9.0 + 2.0)
    v = v + rng.normal(0, 20, size=shape)
    return np.clip(v, 0, 255).astype(np.uint8)


SLICER_SHAPES = [(5, 10, 13), (7, 29, 30), (9, 61, 33), (12, 31, 32), (33, 64, 35), (6, 25, 27)]


def make_slicer():
    digests = {}
    full = {}
    for si, shape in enumerate(SLICER_SHAPES):
        vol = synth_volume(shape, 100 + si)
        for d in range(12):
            ref = po.slicer_oracle(vol, d)
            digests[f"{shape}|{d}"] = {
                "shape": list(ref.shape),
                "sha256_bf16": hashlib.sha256(act_bits(ref, torch.bfloat16).tobytes()).hexdigest(),
                "sha256_f16": hashlib.sha256(act_bits(ref, torch.float16).tobytes()).hexdigest(),
            }
            if si == 0:
                full[f"d{d}"] = ref.astype(np.float32)
    (OUT / "slicer_digests.json").write_text(json.dumps(digests, indent=1))
    np.savez_compressed(OUT / "slicer_small.npz", **full)


def make_merge():
    shape = (6, 7, 9)
    rng = np.random.default_rng(7)
    probs, labels = {}, {}
    # a small palette with engineered fp16 ties: 0.9999 and 0.99995 both round to 1.0;
    # 0.50001 / 0.5 / 0.49999 collide or straddle a half-precision step
    palette = np.array([0.9999, 0.99995, 1.0, 0.5, 0.50001, 0.49999, 0.25, 0.7501, 0.75, 0.3333], np.float32)
    for d in range(12):
        sl = po.direction_slices(np.zeros(shape, np.uint8), d)
        probs[d] = palette[rng.integers(0, len(palette), size=sl.shape)]
        labels[d] = rng.integers(0, 6, size=sl.shape).astype(np.uint8)
    out = {}
    for name, dirs in {"low": [0], "lowY": [1], "lowX": [2], "medium": [0, 1, 2], "high": list(range(12)),
                       "high_nodup": [0, 1, 2, 4, 5, 7, 8, 11]}.items():
        lab, prb = po.merge_injected_oracle(shape, dirs, probs, labels)
        out[f"{name}_labels"] = lab
        out[f"{name}_probs"] = prb.view(np.uint16)
    for d in range(12):
        out[f"in_probs_{d}"] = probs[d]
        out[f"in_labels_{d}"] = labels[d]
    np.savez_compressed(OUT / "merge_injected.npz", **out)


def make_network():
    torch.set_num_threads(8)
    out = {}
    for arch, enc, c in [("unet", "resnet34", 4), ("unetplusplus", "resnext50_32x4d", 6),
                         ("deeplabv3plus", "resnet50", 4), ("unet", "resnet50", 2)]:
        model = make_random_model(arch, enc, c, seed=0)
        vol = structured_volume((2, 40, 70), 11)
        imgs = np.stack([po.preprocess_slice(vol[i]) for i in range(2)]).astype(np.float32)
        with torch.no_grad():
            logits = model(torch.from_numpy(imgs)[:, None]).numpy()
        out[f"{arch}|{enc}|{c}"] = logits.astype(np.float32)
    np.savez_compressed(OUT / "network_logits.npz", **out)


def make_e2e():
    torch.set_num_threads(8)
    model = make_random_model("unet", "resnet34", 4, seed=0)
    pred = po.OraclePredictor(model, 4, batch_size=4)
    vol = structured_volume((20, 40, 45), 21)
    out = {"volume": vol}
    lab, prb = pred.predict_single_axis(vol, True, po.AXIS_Y)
    out["low_y_labels"], out["low_y_probs"] = np.ascontiguousarray(lab), np.ascontiguousarray(prb).view(np.uint16)
    lab, prb = pred.predict_3_ways_max_probs(vol)
    out["medium_labels"], out["medium_probs"] = lab.copy(), prb.copy().view(np.uint16)
    lab, prb = pred.predict_12_ways_max_probs(vol)
    out["high_labels"], out["high_probs"] = lab.copy(), prb.copy().view(np.uint16)
    out["high_one_hot"] = pred.predict_12_ways_one_hot(vol)
    # full per-direction class probabilities (slice space) for the margin clause
    for d in (0, 1, 2):
        _, _, full = pred.predict_single_axis(vol, True, d, return_full=True)
        out[f"full_probs_d{d}"] = full.astype(np.float32)
    np.savez_compressed(OUT / "e2e_unet_r34.npz", **out)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    make_slicer()
    make_merge()
    make_network()
    make_e2e()
    for p in sorted(OUT.iterdir()):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
