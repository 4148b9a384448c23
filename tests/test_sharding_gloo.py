"""world_size-2 gloo test of the N>1 path's host logic: each rank merges its own
work items into packed keys, one all_reduce(MAX) combines them, the result is
bit-identical to the single-process merge."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, out_dir):
    sys.path.insert(0, str(ROOT))
    from oracle import predict_oracle as po
    from volume_segmantics_b200 import sharding

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)  # same data on every rank (replicated volume)
    dirs = sharding.direction_list((1 << 12) - 1, skip_duplicates=False)
    probs = {d: rng.random(sharding.direction_dims(shape, d)).astype(np.float32).round(2) for d in dirs}
    labels = {d: rng.integers(0, 4, sharding.direction_dims(shape, d)).astype(np.uint8) for d in dirs}
    keys = np.zeros(shape, np.uint64)
    for it in sharding.partition(shape, dirs, world, granule=1)[rank]:
        p = np.zeros_like(probs[it.d])
        p[it.s0:it.s1] = probs[it.d][it.s0:it.s1]
        k = sharding.pack_keys_np(probs[it.d], labels[it.d], it.d)
        mask = np.zeros(k.shape, bool)
        mask[it.s0:it.s1] = True
        k = np.where(mask, k, np.uint64(0))
        keys = np.maximum(keys, po.direction_to_volume(k, it.d))
    t = torch.from_numpy(keys.view(np.int64).copy())
    sharding.allreduce_max_keys(t)
    if rank == 0:
        lab, prb = sharding.unpack_keys_np(t.numpy().view(np.uint64))
        want_l, want_p = po.merge_injected_oracle(shape, dirs, probs, labels)
        np.save(Path(out_dir) / "ok.npy", np.array([np.array_equal(lab, want_l), np.array_equal(prb, want_p)]))
    dist.destroy_process_group()


def test_two_rank_key_allreduce_equals_single_merge(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), (6, 9, 11), str(tmp_path)), nprocs=world, join=True)
    ok = np.load(tmp_path / "ok.npy")
    assert ok.all()
