"""Multi-GPU exchange check (run by hand: gpurun --gpus 2 -- python -m torch.distributed.run
--nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py).  The sharded prediction
with (a) the fused NVLink peer reduce+unpack and (b) the NCCL max all-reduce must both equal
the single-GPU prediction bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from volume_segmantics_b200 import sharding  # noqa: E402
from volume_segmantics_b200.engine import Engine  # noqa: E402
from volume_segmantics_b200.plan import B200SegmentationModel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = B200SegmentationModel("U_NET", "resnet34", 4)
shape = (45, 96, 83)  # odd voxel count, ragged dims
vol = np.random.default_rng(9).integers(0, 256, shape, dtype=np.uint8)
nvox = vol.size
eng = Engine(local)
stream = torch.cuda.Stream(device=dev)
eng.set_stream(stream.cuda_stream)
eng.load_model(model)
eng.set_volume(vol)
dirs = sharding.direction_list(0xFFF, True)

# reference: every rank computes everything alone
with torch.cuda.stream(stream):
    eng.reset()
    eng.predict(0xFFF, True)
ref_l, ref_p = eng.fetch()

items = sharding.partition(shape, dirs, world, granule=1)[rank]
shards = sharding.voxel_shards(nvox, world)
per = shards[0][1] - shards[0][0]
v0, v1 = shards[rank]
ok = True

# (a) fused peer exchange
handles = [None] * world
dist.all_gather_object(handles, eng.keys_ipc_handle())
eng.open_peers(handles, rank)
lab = torch.zeros(per, dtype=torch.uint8, device=dev)
prb = torch.zeros(per, dtype=torch.float16, device=dev)
tick = torch.zeros(1, dtype=torch.int32, device=dev)
with torch.cuda.stream(stream):
    eng.reset()
    for it in items:
        eng.predict_range(it.d, it.s0, it.s1)
    dist.all_reduce(tick)
    eng.reduce_unpack_shard(v0, v1, lab.data_ptr(), prb.data_ptr())
    dist.all_reduce(tick)
torch.cuda.synchronize()
got_l, got_p = lab.cpu().numpy()[: v1 - v0], prb.cpu().numpy()[: v1 - v0]
a_ok = np.array_equal(got_l, ref_l.ravel()[v0:v1]) and np.array_equal(got_p.view(np.uint16), ref_p.ravel()[v0:v1].view(np.uint16))
print(f"[rank {rank}] peer reduce+unpack shard [{v0},{v1}): {'OK' if a_ok else 'MISMATCH'}", flush=True)
ok &= a_ok
eng.close_peers()

# (b) NCCL all-reduce
keys = torch.zeros(nvox, dtype=torch.int64, device=dev)
eng.bind_keys(keys.data_ptr())
with torch.cuda.stream(stream):
    keys.zero_()
    for it in items:
        eng.predict_range(it.d, it.s0, it.s1)
    dist.all_reduce(keys, op=dist.ReduceOp.MAX)
torch.cuda.synchronize()
nl, np_ = eng.fetch()
b_ok = np.array_equal(nl, ref_l) and np.array_equal(np_.view(np.uint16), ref_p.view(np.uint16))
print(f"[rank {rank}] NCCL max all-reduce: {'OK' if b_ok else 'MISMATCH'}", flush=True)
ok &= b_ok
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU CHECK", "PASSED" if int(flag.item()) else "FAILED", flush=True)
dist.destroy_process_group()
