"""Build libvsb200.so (sm_100a only) in-tree with nvcc.

Usage: python -m volume_segmantics_b200.build [--force]
The shared object lands next to this file so it travels to the GPU box with
the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libvsb200.so"          # fp16 activations/weights (default)
LIB_BF16 = HERE / "libvsb200_bf16.so"  # bfloat16 variant (VSB200_VARIANT=bf16)
SOURCES = ["engine.cu", "conv_tc.cu", "conv_halo.cu", "conv_stem.cu", "kernels_simple.cu", "kernels_head.cu", "kernels_ingest.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _stale() -> bool:
    if not LIB.exists() or not LIB_BF16.exists():
        return True
    t = min(LIB.stat().st_mtime, LIB_BF16.stat().st_mtime)
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(HERE.parent / "include" / "vsb200.h")
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = True) -> Path:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    builddir = HERE / "build"
    builddir.mkdir(exist_ok=True)
    variants = [("f16", LIB, "-DVSB_ACT_F16=1"), ("bf16", LIB_BF16, "-DVSB_ACT_F16=0")]
    procs = []
    for tag, _, define in variants:
        for src in SOURCES:
            obj = builddir / f"{src}.{tag}.o"
            cmd = [nvcc, *NVCC_FLAGS, define, "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    for tag, lib, _ in variants:
        objs = [str(builddir / f"{src}.{tag}.o") for src in SOURCES]
        # the CUDA runtime is linked as a shared library (libcudart.so.12, the one torch has already
        # loaded when the host shim runs): the shipped artefact then carries none of the runtime's own
        # entry-point names, and the engine shares one runtime instance with the plumbing
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "shared",
               "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", str(lib), *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
