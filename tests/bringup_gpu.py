"""GPU bring-up diagnostics (run by hand under gpurun, not collected by pytest):
per-convolution tcgen05-vs-CUDA-core cross-checks, then whole-network and
end-to-end comparisons against the CPU oracle.  Prints, never asserts."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import predict_oracle as po  # noqa: E402
from oracle.smp_models import make_random_model  # noqa: E402
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.netspec import NetSpec  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel  # noqa: E402


class _Probe(B200SegmentationModel):
    """A hand-built NetSpec with random parameters (unit tests of single convs)."""

    def __init__(self, spec, seed=0):
        torch.nn.Module.__init__(self)
        self.spec, self.classes, self._engine = spec, spec.classes, None
        g = torch.Generator().manual_seed(seed)
        from volume_segmantics_b200.plan import _attach
        for key, (shape, (role, _)) in spec.param_shapes().items():
            if role == "conv_w":
                fan = shape[1] * shape[2] * shape[3]
                _attach(self, key, torch.randn(shape, generator=g) * (2.0 / fan) ** 0.5, True)
            elif role == "conv_b":
                _attach(self, key, torch.randn(shape, generator=g) * 0.1, True)
            elif role in ("bn_w", "bn_var"):
                _attach(self, key, torch.rand(shape, generator=g) + 0.5, role == "bn_w")
            elif role in ("bn_b", "bn_mean"):
                _attach(self, key, torch.randn(shape, generator=g) * 0.1, role == "bn_b")
            else:
                _attach(self, key, torch.tensor(0, dtype=torch.long), False)


def probe_spec(cin, cout, k=3, stride=1, dil=1, up=False, skip=0, res=False, relu=True, f32=False):
    """1-channel input -> (CUDA-core conv to `cin` channels at the right scale)
    -> conv under test -> 4-class head."""
    net = NetSpec("probe", "none", 4)
    lo = net.conv("a", [0], cin, 3, pad=1, bias=True, relu=True)
    srcs = [(lo, 0)]
    if up:
        srcs = [(net.conv("lo", [lo], cin, 3, stride=2, pad=1, bias=True, relu=True), 1)]
        if skip:
            srcs.append((net.conv("sk", [0], skip, 3, pad=1, bias=True, relu=True), 0))
    r = -1
    pad = dil * (k // 2)
    if res:
        r = net.conv("r", [0], cout, 3, stride=stride, pad=1, bias=True)
    t = net.conv("t", srcs, cout, k, stride=stride, pad=pad, dil=dil, bn="tbn", relu=relu, res=r,
                 out_dtype=1 if f32 else 0)
    if f32:
        net.head(t)
        return net, t
    hd = net.conv("h", [t], 4, 3, pad=1, bias=True, out_dtype=1)
    net.head(hd)
    return net, t


def run_probe(eng, name, hp, wp, nb, **kw):
    spec, t = probe_spec(**kw)
    if kw.get("f32"):
        spec.classes = kw["cout"]
    model = _Probe(spec)
    rng = np.random.default_rng(1)
    imgs = rng.standard_normal((nb, hp, wp)).astype(np.float32)
    outs = {}
    for impl in ("generic", "tc"):
        eng.set_conv_impl(impl)
        try:
            eng.forward_logits(model, imgs)
            outs[impl] = eng.debug_tensor(t)
        except Exception as ex:  # noqa: BLE001
            print(f"[probe] {name:34s} {impl}: EXCEPTION {ex}")
            return False
    a, b = outs["generic"], outs["tc"]
    err = np.abs(a - b)
    scale = np.abs(a).max() + 1e-9
    bad = int((err > 0.02 * scale + 0.02).sum())
    print(f"[probe] {name:34s} shape={a.shape} max|ref|={scale:.3f} maxerr={err.max():.4f} "
          f"meanerr={err.mean():.5f} bad={bad}/{a.size} {'OK' if bad == 0 else 'MISMATCH'}")
    if bad:
        idx = np.argwhere(err > 0.02 * scale + 0.02)[:6]
        for i in idx:
            print("         at", tuple(i), "ref", a[tuple(i)], "tc", b[tuple(i)])
    return bad == 0


def main():
    t0 = time.time()
    eng = Engine(0)
    print("engine created", time.time() - t0)
    ok = True
    probes = [
        ("3x3 64->64 relu", dict(cin=64, cout=64)),
        ("3x3 64->64 res", dict(cin=64, cout=64, res=True)),
        ("3x3 128->256", dict(cin=128, cout=256)),
        ("3x3 32->32 (SW64)", dict(cin=32, cout=32)),
        ("3x3 16->16 (SW32)", dict(cin=16, cout=16)),
        ("3x3 16->4 f32 head", dict(cin=16, cout=4, f32=True, relu=False)),
        ("1x1 256->512 (2 n-tiles)", dict(cin=256, cout=512, k=1)),
        ("3x3 s2 64->128", dict(cin=64, cout=128, stride=2)),
        ("1x1 s2 64->128 norelu", dict(cin=64, cout=128, k=1, stride=2, relu=False)),
        ("3x3 dil2 64->64", dict(cin=64, cout=64, dil=2)),
        ("PS up(64)+skip(64)->32", dict(cin=64, cout=32, up=True, skip=64)),
        ("PS up(32)->16", dict(cin=32, cout=16, up=True)),
        ("PS up(128)+skip(64)->256", dict(cin=128, cout=256, up=True, skip=64)),
        ("3x3 48->48 (KB16 x3)", dict(cin=48, cout=48)),
    ]
    allp = [(n, 64, 96, 3, kw) for n, kw in probes]
    allp += [(n + " [32x32,nb5]", 32, 32, 5, kw) for n, kw in probes[:3] + probes[7:8] + probes[10:11]]
    allp += [(n + " [160x224,nb2]", 160, 224, 2, kw) for n, kw in probes[:1] + probes[10:11]]
    args = sys.argv[1:]
    if args and args[0] == "--count":
        print(len(allp))
        return
    if args and args[0] == "--probe":
        for i in args[1].split(","):
            n, hp, wp, nb, kw = allp[int(i)]
            ok &= run_probe(eng, f"#{i} " + n, hp, wp, nb, **kw)
        print("probes", "ALL OK" if ok else "SOME FAILED", time.time() - t0)
        return

    # ---- whole network vs oracle ------------------------------------------------
    torch.set_num_threads(8)
    for arch, mt, enc, C in [("unet", "U_NET", "resnet34", 4)]:
        oracle = make_random_model(arch, enc, C, seed=0)
        model = B200SegmentationModel(mt, enc, C)
        model.load_state_dict(oracle.state_dict())
        rng = np.random.default_rng(2)
        vol = rng.integers(0, 256, size=(4, 64, 96), dtype=np.uint8)
        imgs = np.stack([po.preprocess_slice(vol[i]) for i in range(vol.shape[0])]).astype(np.float32)
        with torch.no_grad():
            ref = oracle(torch.from_numpy(imgs)[:, None]).permute(0, 2, 3, 1).numpy()
        for impl in ("generic", "simt", "tc"):
            eng.set_conv_impl(impl)
            try:
                out = eng.forward_logits(model, imgs)
            except Exception as ex:  # noqa: BLE001
                print(f"[net] {arch} {impl}: EXCEPTION {ex}")
                continue
            err = np.abs(out - ref)
            pr = torch.softmax(torch.from_numpy(ref), -1).numpy()
            po_ = torch.softmax(torch.from_numpy(out), -1).numpy()
            print(f"[net] {arch}/{enc} {impl}: logits max|ref|={np.abs(ref).max():.3f} maxerr={err.max():.4f} "
                  f"meanerr={err.mean():.5f} prob maxerr={np.abs(pr - po_).max():.4f} "
                  f"label agree={(pr.argmax(-1) == po_.argmax(-1)).mean():.5f}")
    eng.set_conv_impl("tc")
    print("done", time.time() - t0)


if __name__ == "__main__":
    main()
