"""Multi-GPU sharding of one prediction.

Every (direction, slice) is an independent forward pass; the only coupling is
the per-voxel maximum over directions.  So: the uint8 volume is replicated,
the (direction, slice-range) work items are split across ranks by padded pixel
count, every rank merges its own items into a full-size packed-key volume, and
ONE ``all_reduce(MAX)`` over the keys (NCCL over NVLink; gloo in the CPU tests)
combines them.  Max over packed keys is associative and commutative, so the
result is bit-identical to the single-GPU merge.

The reference has no multi-device path at all (one process, ``cuda_device``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib

DUPLICATE_OF = {3: 1, 6: 4, 9: 7, 10: 0}


@dataclass(frozen=True)
class WorkItem:
    d: int
    s0: int
    s1: int
    cost: int  # padded pixels

    @property
    def slices(self) -> int:
        return self.s1 - self.s0


def direction_list(dir_mask: int, skip_duplicates: bool = True) -> List[int]:
    dirs = [d for d in range(12) if dir_mask >> d & 1]
    if skip_duplicates:
        dirs = [d for d in dirs if not (d in DUPLICATE_OF and DUPLICATE_OF[d] in dirs)]
    return dirs


def direction_dims(shape_zyx: Sequence[int], d: int) -> Tuple[int, int, int]:
    """(S, H, W) of direction d = 3k + a, from the reference's rot90 + swapaxes."""
    z, y, x = shape_zyx
    k, a = divmod(d, 3)
    ni, nj = (y, z) if k & 1 else (z, y)
    return {0: (ni, nj, x), 1: (nj, ni, x), 2: (x, nj, ni)}[a]


def _pad32(v: int) -> int:
    return (v + 31) // 32 * 32


def partition(shape_zyx: Sequence[int], dirs: Sequence[int], world: int, granule: int = 8) -> List[List[WorkItem]]:
    """Contiguous split of the concatenated slice sequence of all directions
    into `world` shares of (nearly) equal padded-pixel cost, cut on multiples of
    `granule` slices where possible so batches stay whole."""
    spans = []
    total = 0
    for d in dirs:
        s, h, w = direction_dims(shape_zyx, d)
        c = _pad32(h) * _pad32(w)
        spans.append((d, s, c))
        total += s * c
    shares: List[List[WorkItem]] = [[] for _ in range(world)]
    rank, acc = 0, 0
    for d, s, c in spans:
        pos = 0
        while pos < s:
            target = (total * (rank + 1) + world - 1) // world
            room = max(0, target - acc)
            n = min(s - pos, max(1, -(-room // c)))
            if pos + n < s and n > granule:
                n = n // granule * granule
            if rank == world - 1:
                n = s - pos
            shares[rank].append(WorkItem(d, pos, pos + n, n * c))
            pos += n
            acc += n * c
            if acc >= target and rank < world - 1:
                rank += 1
    return shares


def voxel_shards(nvox: int, world: int, align: int = 8) -> List[Tuple[int, int]]:
    """Contiguous voxel ranges, one per rank, for the fused peer reduce + unpack
    (cut on multiples of `align` voxels so 16-byte key loads stay aligned)."""
    per = -(-nvox // world)
    per = -(-per // align) * align
    return [(min(r * per, nvox), min((r + 1) * per, nvox)) for r in range(world)]


# ---- packed keys on the host (tests, diagnostics) -----------------------------
def pack_keys_np(prob_f32: np.ndarray, labels: np.ndarray, d: int) -> np.ndarray:
    """Host restatement of the device key layout (csrc/kernels.h pack_key)."""
    h = prob_f32.astype(np.float16).view(np.uint16).astype(np.uint64)
    f = prob_f32.astype(np.float32).view(np.uint32).astype(np.uint64)
    return (h << np.uint64(48)) | (np.uint64(15 - d) << np.uint64(44)) | (
        labels.astype(np.uint64) << np.uint64(36)) | f


def unpack_keys_np(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    keys = keys.astype(np.uint64)
    labels = ((keys >> np.uint64(36)) & np.uint64(0xFF)).astype(np.uint8)
    probs = (keys >> np.uint64(48)).astype(np.uint16).view(np.float16)
    return labels, probs


def allreduce_max_keys(keys_i64) -> None:
    """In-place MAX all-reduce of a torch int64 view of the key volume.  Keys are
    < 2^63 (fp16 bits of a non-negative probability lead), so signed and
    unsigned order agree."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(keys_i64, op=dist.ReduceOp.MAX)
