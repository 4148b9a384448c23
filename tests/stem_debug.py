"""Bring-up aid for the v3 stem (run under gpurun): each experiment in its own process."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from volume_segmantics_b200.engine import Engine
from volume_segmantics_b200.plan import B200SegmentationModel
import torch
torch.manual_seed(0)
dbg, ver = int(sys.argv[1]), int(sys.argv[2])
m = B200SegmentationModel("U_NET", "resnet34", 4)
e = Engine(0)
e.set_flag("stem", ver)
e.set_flag("stem_dbg", dbg)
e.set_flag("sync_each", 1)
x = np.random.default_rng(0).normal(0, 1, (2, 96, 128)).astype(np.float32)
out = e.forward_logits(m, x)
print("OK dbg", dbg, "ver", ver, float(np.abs(out).max()))
if dbg in (0, 8):
    e.set_flag("stem", 2)
    e.set_flag("stem_dbg", 0)
    ref = e.forward_logits(m, x)
    print("   max |v3 - v2| =", float(np.abs(out - ref).max()), "bit-identical" if np.array_equal(out, ref) else "DIFFERENT")
    # where do the stem outputs differ?  tensor 1 = conv output (H/2), tensor 2 = pooled (H/4)
    for tid in (1, 2):
        e.set_flag("stem", 3); e.forward_logits(m, x); a = e.debug_tensor(tid)
        e.set_flag("stem", 2); e.forward_logits(m, x); b = e.debug_tensor(tid)
        d = np.abs(a - b).max(axis=(0, 3))
        ys, xs = np.nonzero(d > 0)
        print("   tensor", tid, a.shape, "differing pixels", len(ys), "rows", sorted(set(ys.tolist()))[:40], "cols", sorted(set(xs.tolist()))[:64])
''' % str(ROOT)
for dbg, ver in [(0, 3)]:
    r = subprocess.run([sys.executable, "-c", CHILD, str(dbg), str(ver)], capture_output=True, text=True, timeout=120)
    tail = (r.stdout.strip().splitlines() or ["<no stdout>"])
    err = [l for l in r.stderr.strip().splitlines() if "Error" in l or "error" in l][-1:] if r.returncode else []
    print(f"dbg={dbg} ver={ver} rc={r.returncode}", " | ".join(tail), " | ".join(err), flush=True)
