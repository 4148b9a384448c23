"""Per-kernel SASS evidence of libvsb200.so (run here, no GPU needed):
    python tests/sass_summary.py > profiles/r02_sass_summary.txt
Counts, from `cuobjdump -sass`, the instructions that prove which hardware path a kernel uses:
UTCHMMA (tcgen05.mma kind::f16), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (cp.async.bulk 1-D), UTCBAR (tcgen05.commit), LDGSTS (cp.async), SYNCS (mbarrier),
RED / ATOM(G) (global reductions / atomics)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = ROOT / "volume_segmantics_b200" / (sys.argv[1] if len(sys.argv) > 1 else "libvsb200.so")
txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDGSTS", "SYNCS", "RED", "ATOM"]
rows = []
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    mangled = part.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"^void ", "", dem)
    depth, cut = 0, len(dem)
    for i in range(len(dem) - 1, -1, -1):  # strip the trailing parameter list only
        if dem[i] == ")":
            depth += 1
        elif dem[i] == "(":
            depth -= 1
            if depth == 0:
                cut = i
                break
    dem = dem[:cut].replace("(anonymous namespace)::", "")
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", part)
    c = collections.Counter()
    for op in ops:
        base = "ATOM" if op.startswith("ATOM") else ("RED" if op in ("RED", "REDG", "REDUX") else op)
        c[base] += 1
    rows.append((dem, len(ops), c))
print(f"# {lib.name}: instruction counts per kernel from `cuobjdump -sass` (see tests/sass_summary.py)")
print(f"{'kernel':72s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in keys))
for dem, n, c in sorted(rows):
    print(f"{dem[:72]:72s} {n:6d} " + " ".join(f"{c[k]:7d}" for k in keys))
