"""Several GPUs behind ONE ``VolSeg2dPredictor`` (SURVEY.md 5 "config" row and 8e).

The reference binds one device (``settings.cuda_device``,
vol_seg_2d_predictor.py:22).  The additive settings key ``cuda_devices: [0, 1, ...]``
(absent => the reference's behaviour) makes the drop-in drive one libvsb200 engine per
listed GPU from this single host process:

  1. ingest   every GPU uploads 1/N of the uint8 volume over its own PCIe link and pulls the
              other parts from its peers over NVLink (``vsb_set_volume_shard`` +
              ``vsb_volume_pull``), instead of N uploads of the whole volume;
  2. predict  the (direction, slice-range) work items are split by padded-pixel cost
              (``sharding.partition``); every engine merges its items into its own packed-key
              volume -- no communication;
  3. exchange ONE fused kernel per GPU max-reduces its voxel shard over all peers' key
              volumes through NVLink peer loads and unpacks it (``vsb_fetch_shard``), and each
              GPU downloads its own shard of labels / probabilities over its own PCIe link
              into the shared page-locked result.

Max over packed keys is associative and commutative, so the result equals the single-GPU
prediction bit for bit (tests/test_multi_gpu.py).  One Python thread per engine issues the
(asynchronous) C-ABI calls; ctypes releases the GIL for their duration.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import sharding
from .engine import Engine, get_engine
from .plan import B200SegmentationModel


class LocalGroup:
    """N engines of this process working on one volume."""

    def __init__(self, devices: Sequence[int]):
        devices = [int(d) for d in devices]
        if len(devices) < 2 or len(set(devices)) != len(devices) or len(devices) > 8:
            raise ValueError(f"cuda_devices must list 2..8 distinct GPUs, got {devices}")
        self.devices = devices
        self.engines: List[Engine] = [get_engine(d) for d in devices]
        self._pool = ThreadPoolExecutor(max_workers=len(devices), thread_name_prefix="vsb200-gpu")

    def predict(self, model: B200SegmentationModel, vol: np.ndarray, dir_mask: int, want_probs: bool = True,
                skip_duplicates: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        if vol.dtype != np.uint8 or vol.ndim != 3:
            raise ValueError("the multi-GPU path ingests 3-D uint8 volumes")
        vol = np.ascontiguousarray(vol)
        n = len(self.engines)
        shape = tuple(int(v) for v in vol.shape)
        nvox = vol.size
        dirs = sharding.direction_list(dir_mask, skip_duplicates)
        parts = sharding.partition(shape, dirs, n, granule=8)
        shards = sharding.voxel_shards(nvox, n)
        lead = self.engines[0]
        labels = lead._host_buffer(shape, np.uint8)
        probs = lead._host_buffer(shape, np.float16) if want_probs else None
        labels_flat = labels.reshape(-1)
        probs_flat = probs.reshape(-1) if want_probs else None
        barrier = threading.Barrier(n)

        def work(r: int) -> None:
            eng = self.engines[r]
            try:
                eng.ensure_model(model)
                eng.set_vote_mode(False)
                eng.set_volume_shard(vol, *shards[r])  # allocates the whole volume + zeroed keys
                eng.synchronize()
                barrier.wait()  # every part is in some GPU's HBM, every key volume exists
                eng.attach_peers(self.engines, r)  # enables peer access: the pulls below go over NVLink
                for q in range(n):
                    if q != r:
                        eng.volume_pull(self.engines[q], *shards[q])
                for it in parts[r]:
                    eng.predict_range(it.d, it.s0, it.s1)
                eng.synchronize()
                barrier.wait()  # every engine's keys are final
                eng.fetch_shard(shards[r][0], shards[r][1], labels_flat, probs_flat)
                barrier.wait()  # every shard is fetched: the key volumes may be reused
            except BaseException:
                barrier.abort()
                raise

        futures = [self._pool.submit(work, r) for r in range(n)]
        errors = []
        for f in futures:
            try:
                f.result()
            except threading.BrokenBarrierError:
                pass
            except BaseException as ex:  # noqa: BLE001
                errors.append(ex)
        if errors:
            raise errors[0]
        return labels, probs

    def close(self) -> None:
        self._pool.shutdown(wait=True)
