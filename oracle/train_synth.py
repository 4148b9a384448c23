"""A few seconds of CPU training so the parity network is *decisive* like a real
checkpoint (TEST INFRASTRUCTURE ONLY).

Random-init weights give softmax outputs within 0.004 of uniform (measured: p_max
in 0.254..0.258 for 4 classes), so every voxel sits on a decision boundary and
label agreement measures nothing but rounding noise.  The reference's users run
trained models (vol_seg_2d_trainer.py); this restates just enough of that --
cross-entropy on synthetic labels derived from the synthetic volume -- to move
the logits away from the boundaries.  Deterministic for a fixed torch build and
thread count; tests compare engine and oracle on the SAME in-memory weights, so
cross-machine reproducibility of the training itself is not required.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import predict_oracle as po
from .make_golden import structured_volume
from .smp_models import OracleSegModel


def synthetic_labels(vol: np.ndarray, classes: int) -> np.ndarray:
    """Smooth the volume (box filter) and quantise into `classes` intensity bands."""
    v = torch.from_numpy(vol.astype(np.float32))[None, None]
    v = F.avg_pool3d(F.pad(v, (2, 2, 2, 2, 2, 2), mode="replicate"), 5, stride=1)[0, 0].numpy()
    qs = np.quantile(v, np.linspace(0, 1, classes + 1)[1:-1])
    return np.digitize(v, qs).astype(np.int64)


def make_trained_model(arch="unet", encoder="resnet34", classes=4, seed=0, steps=60, size=64):
    torch.manual_seed(seed)
    model = OracleSegModel(arch, encoder, classes)
    vol = structured_volume((size, size, size), 31 + seed)
    lab = synthetic_labels(vol, classes)
    rng = np.random.default_rng(seed)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    model.train()
    for _ in range(steps):
        xs, ys = [], []
        for _ in range(8):
            a = int(rng.integers(0, 3))
            i = int(rng.integers(0, size))
            xs.append(po.preprocess_slice(np.take(vol, i, axis=a)))
            ys.append(np.take(lab, i, axis=a))
        x = torch.from_numpy(np.stack(xs).astype(np.float32))[:, None]
        y = torch.from_numpy(np.stack(ys))
        loss = F.cross_entropy(model(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    return model.eval(), float(loss.detach())
