#!/usr/bin/env python
"""Benchmark of the prediction hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric: voxels/sec of a 12-direction ('high' quality) U-Net/ResNet-34, 4-class
prediction of a synthetic 1024^3 uint8 volume (BASELINE.json configs[2], the
configuration the metric is quoted on; it fits one B200).  One "step" = one
full prediction of the volume.  The same volume is sharded over N GPUs
(strong scaling): work items = (direction, slice range), one NCCL max-reduce of
the packed keys, rank 0 unpacks.

  value : volume already resident in HBM -> label + fp16 probability volumes in
          HBM (CUDA events on the engine's stream, max over ranks).
  e2e   : VolSeg2dPredictor._predict_12_ways_max_probs(host ndarray) -> host
          ndarrays; pinned H2D of the volume and D2H of labels + probs inside
          the timed region (N = 1: through the reference-facing API itself).
  roofline : the tcgen05 convolution kernel; achieved = algorithmic conv FLOPs
          of the launches / their summed CUDA-event durations.
  cpu_baseline : the CPU oracle (restatement of the reference path) on a bounded
          sample, extrapolated; a reported baseline, not a target.

--impl reference runs ONLY the CPU oracle port (the reference itself cannot be
imported in this image: h5py / segmentation_models_pytorch / albumentations are
absent) with all host threads on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ARCH, ENCODER, CLASSES = "U_NET", "resnet34", 4
DIR_MASK = (1 << 12) - 1
METRIC = "voxels/sec, 12-direction U-Net prediction of a 1024^3 volume"


def env_int(name, default):
    return int(os.environ.get(name, default))


def synth_volume(size):
    # integers(0, 256) uint8 volume, fixed seed
    return np.random.default_rng(20240).integers(0, 256, size=(size, size, size), dtype=np.uint8)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.path = None, f"/tmp/vsb_clocks_{os.getpid()}.csv"
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) == 6 and parts[0].isdigit():
                rows.append(parts)
        if rows:
            sm = sorted(int(r[0]) for r in rows)
            busy = [v for v in sm if v > 0.5 * sm[-1]] or sm
            out["sm_mhz"] = busy[len(busy) // 2]
            out["sm_max_mhz"] = int(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return out


# ----------------------------------------------------------------------------- CPU oracle leg
def cpu_oracle_sample(size, slices, threads):
    """Time the oracle on `slices` Z-slices of a size^2 image + one merge sample;
    extrapolate to the full 12-direction prediction.  Returns (voxels/s, detail)."""
    import torch

    from oracle import predict_oracle as po
    from oracle.smp_models import make_random_model

    torch.set_num_threads(threads)
    model = make_random_model("unet", ENCODER, CLASSES, seed=0)
    pred = po.OraclePredictor(model, CLASSES, batch_size=4)
    vol = np.random.default_rng(1).integers(0, 256, size=(slices, size, size), dtype=np.uint8)
    t0 = time.perf_counter()
    pred.predict_single_axis(vol, True, po.AXIS_Z)
    t_slice = (time.perf_counter() - t0) / slices
    mz = max(1, min(size, (32 << 20) // (size * size)))
    pc = np.random.default_rng(2).random((2, mz, size, size)).astype(np.float16)
    lc = np.zeros((2, mz, size, size), np.uint8)
    t0 = time.perf_counter()
    po.merge_vols_in_mem(pc, lc)
    t_merge_vox = (time.perf_counter() - t0) / (mz * size * size)
    nvox = size ** 3
    total = 12 * size * t_slice + 11 * nvox * t_merge_vox
    return nvox / total, {"s_per_slice": t_slice, "s_per_merge_voxel": t_merge_vox}


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    size = args.size
    slices = 4
    vals = []
    for i in range(args.warmup + args.steps):
        v, _ = cpu_oracle_sample(size, slices, threads)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    sample = (f"{slices} Z-slices of {size}x{size} through the fp32 CPU oracle (batch 4) + one "
              f"(2,{max(1, min(size, (32 << 20) // (size * size)))},{size},{size}) fp16 merge per step, "
              f"extrapolated linearly to 12 directions x {size} slices + 11 merges")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * size ** 3 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"U-Net/ResNet-34, high quality (12 directions), {CLASSES} classes, synthetic {size}^3 uint8 volume",
                   "weights": "random init (seed 0), BN statistics randomised"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU leg
def run_ours(args):
    import torch
    import torch.distributed as dist

    from volume_segmantics_b200 import _lib, sharding
    from volume_segmantics_b200.engine import Engine
    from volume_segmantics_b200.plan import B200SegmentationModel, conv_macs_per_pixel

    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    size = args.size
    nvox = size ** 3

    torch.manual_seed(0)
    model = B200SegmentationModel(ARCH, ENCODER, CLASSES)  # random init of the named architecture
    g = torch.Generator().manual_seed(1)
    sd = model.state_dict()
    for key, (shape, (role, _)) in model.spec.param_shapes().items():  # randomise BN
        if role in ("bn_w", "bn_var"):
            sd[key].copy_(torch.rand(shape, generator=g) + 0.5)
        elif role in ("bn_b", "bn_mean"):
            sd[key].copy_(torch.randn(shape, generator=g) * 0.1)

    vol_host = torch.from_numpy(synth_volume(size)).pin_memory()
    eng = Engine(local)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    eng.load_model(model)
    if args.batch:
        eng.set_batch(args.batch)
    vol_dev = vol_host.to(dev)
    torch.cuda.synchronize()
    eng.set_volume_device(vol_dev.data_ptr(), (size, size, size))
    dirs = sharding.direction_list(DIR_MASK, skip_duplicates=True)
    items = sharding.partition((size, size, size), dirs, world, granule=8)[rank]

    # ---- the one exchange step --------------------------------------------
    # "peer": fused max-reduce + unpack of this rank's voxel shard, reading the other ranks'
    #         key volumes over NVLink through CUDA-IPC mappings (one kernel, no NCCL payload);
    #         two 4-byte all-reduces act as stream-ordered barriers around it.
    # "nccl": ncclAllReduce(max) over the whole 8 B/voxel key volume, then rank 0 unpacks.
    exchange = args.exchange if world > 1 else "none"
    keys = None
    if exchange == "peer":
        handles = [None] * world
        dist.all_gather_object(handles, eng.keys_ipc_handle())
        ok = 1
        try:
            eng.open_peers(handles, rank)
        except _lib.VsbError as ex:
            print(f"[rank {rank}] peer mapping unavailable ({ex}); falling back to the NCCL all-reduce", file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            eng.close_peers()
            exchange = "nccl"
    if exchange == "nccl":
        keys = torch.zeros(nvox, dtype=torch.int64, device=dev)
        eng.bind_keys(keys.data_ptr())
    shards = sharding.voxel_shards(nvox, world)
    v0, v1 = shards[rank]
    out_n = nvox if exchange != "peer" else shards[0][1] - shards[0][0]
    labels_dev = torch.empty(out_n, dtype=torch.uint8, device=dev)
    probs_dev = torch.empty(out_n, dtype=torch.float16, device=dev)
    tick = torch.zeros(1, dtype=torch.int32, device=dev)

    def step():
        with torch.cuda.stream(stream):
            if exchange == "nccl":
                keys.zero_()
            else:
                eng.reset()
            for it in items:
                eng.predict_range(it.d, it.s0, it.s1)
            if exchange == "nccl":
                dist.all_reduce(keys, op=dist.ReduceOp.MAX)
                if rank == 0:
                    eng.unpack_device(labels_dev.data_ptr(), probs_dev.data_ptr())
            elif exchange == "peer":
                dist.all_reduce(tick)  # every rank has finished merging into its own keys
                eng.reduce_unpack_shard(v0, v1, labels_dev.data_ptr(), probs_dev.data_ptr())
                dist.all_reduce(tick)  # every rank has finished reading its peers' keys
            else:
                eng.unpack_device(labels_dev.data_ptr(), probs_dev.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    def timed_pass(profile):
        """K steps bracketed by barrier + synchronize, timed with CUDA events on the engine's stream."""
        eng.set_profiling(profile)
        eng.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
        for _ in range(args.steps):
            step()
        with torch.cuda.stream(stream):
            e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    # pass 1 (the reported value): no per-launch events.  pass 2: the same K steps with every
    # kernel launch bracketed by CUDA events, for the roofline of the conv kernels.
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed_pass(False)
    launches = eng.launch_count()  # own kernels only (NCCL kernels are not counted)
    ms_prof = ms if args.no_profile else timed_pass(True)
    clocks = sampler.stop() if sampler else None
    stages = eng.stage_times()
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = nvox / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 conv), this rank's launches ----
    peak_tf, peak_gbs, peak_src = measured_peaks()
    macs_px = conv_macs_per_pixel(model.spec) - 49 * 64 / 4.0  # the 7x7 stem is timed in its own class
    padded_px = sum(it.cost for it in items)
    conv_ms, conv_n = stages["conv_tc"]
    conv_flops = 2.0 * macs_px * padded_px * args.steps
    roofline = None
    traffic = None  # DRAM bytes per conv launch from the committed ncu capture of one batch (profiles/)
    tpath = ROOT / "profiles" / "r01_conv_traffic.json"
    if tpath.exists() and size == 1024:
        traffic = json.loads(tpath.read_text()).get("dram_bytes_per_conv_launch")
    if conv_ms > 0:
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tcgen05 conv kernels (conv_halo_kernel, conv_halo2_kernel<>, conv_tc_kernel<>)", "achieved": ach,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, average over the conv launches of one "
                                      "batch of 32 slices of 1024^2 (profiles/r01_ncu_batch32_dram.csv)" if traffic else None,
                    "peak_source": peak_src,
                    "launches": conv_n, "avg_launch_ms": conv_ms / max(1, conv_n),
                    "flops_per_launch": conv_flops / max(1, conv_n),
                    "share_of_step": conv_ms / ms_prof,
                    "measured": "second pass of the same K steps with per-launch CUDA events "
                                f"({ms_prof / args.steps:.1f} ms/step with events vs {ms / args.steps:.1f} without)",
                    "other_stage_ms_per_step": {k: v[0] / args.steps for k, v in stages.items() if k != "conv_tc"}}

    # ---- e2e through the public API with host buffers ------------------------------
    e2e = None
    if rank == 0 or world > 1:
        e2e = run_e2e(args, eng, model, vol_host, stream, world, rank, items, keys, dev, exchange, shards)

    # ---- CPU baseline: bounded sample on rank 0, N = 1 only -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, detail = cpu_oracle_sample(size, 8 if size >= 1024 else 16, threads)
        cpu = {"value": v, "unit": "voxels/s", "cores": threads, "kind": "port",
               "sample": f"{8 if size >= 1024 else 16} slices of {size}^2 through the fp32 CPU oracle + one fp16 merge sample, "
                         f"extrapolated to 12 directions x {size} slices + 11 merges ({detail['s_per_slice']:.3f} s/slice, "
                         f"{detail['s_per_merge_voxel'] * 1e9:.1f} ns/merge-voxel)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp16" if _lib.act_dtype() == torch.float16 else "bf16", "data": "synthetic",
            "config": {"workload": f"U-Net/ResNet-34, high quality (12 directions), {CLASSES} classes, synthetic {size}^3 uint8 volume",
                       "weights": "random init of the named architecture (seed 0), BN statistics randomised",
                       "directions_computed": len(dirs),
                       "note": "directions 3,6,9,10 duplicate 1,4,7,0 image-for-image and can never win the first-max merge "
                               "(SURVEY.md 3.3); they are skipped and NOT counted in the roofline FLOPs",
                       "l2": "inputs larger than L2 (1 GiB volume + 8 GiB keys per step)",
                       "accumulate": "fp32", "parallelism": f"slice-range sharding over {world} GPU(s)",
                       "exchange": {"none": "single GPU: unpack only",
                                    "peer": "fused max-reduce + unpack of each rank's voxel shard over NVLink peer memory (CUDA IPC); result sharded over ranks",
                                    "nccl": "ncclAllReduce(max) of the 8 B/voxel key volume, rank 0 unpacks"}[exchange]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, eng, model, vol_host, stream, world, rank, items, keys, dev, exchange, shards):
    """Host ndarray in -> host ndarrays out, copies inside the timed region."""
    import torch
    import torch.distributed as dist

    size = args.size
    nvox = size ** 3
    vol_np = vol_host.numpy()
    steps = max(1, min(args.steps, 2))
    if world == 1:
        # exactly the call a user of the reference makes
        from types import SimpleNamespace

        from volume_segmantics_b200.host.predictor import VolSeg2dPredictor

        pred = VolSeg2dPredictor.__new__(VolSeg2dPredictor)
        pred.settings = SimpleNamespace(cuda_device=eng.device)
        pred.model_device_num, pred.model, pred.num_labels, pred.label_codes = eng.device, model, CLASSES, {}
        pred._engine = eng
        eng.bind_keys(0)
        eng.set_stream(0)
        pred._predict_12_ways_max_probs(vol_np)  # warm-up (allocations)
        labels = probs = None
        t0 = time.perf_counter()
        for _ in range(steps):
            labels = probs = None  # the caller is done with the previous result: its pinned block is reused
            labels, probs = pred._predict_12_ways_max_probs(vol_np)
        dt = (time.perf_counter() - t0) / steps
        assert labels.shape == vol_np.shape and probs.dtype == np.float16
        return {"value": nvox / dt, "unit": "voxels/s", "h2d_bytes_per_step": nvox, "d2h_bytes_per_step": 3 * nvox,
                "api": "VolSeg2dPredictor._predict_12_ways_max_probs(ndarray) -> (uint8, float16) ndarrays"}
    # N > 1: every rank uploads the (replicated) volume, rank 0 downloads the result
    labels_h = torch.empty(nvox, dtype=torch.uint8).pin_memory() if rank == 0 else None
    probs_h = torch.empty(nvox, dtype=torch.float16).pin_memory() if rank == 0 else None
    vol_dev = torch.empty(nvox, dtype=torch.uint8, device=dev)
    per = shards[0][1] - shards[0][0]
    v0, v1 = shards[rank]
    n_out = nvox if exchange == "nccl" else per
    labels_dev = torch.empty(n_out, dtype=torch.uint8, device=dev)
    probs_dev = torch.empty(n_out, dtype=torch.float16, device=dev)
    lab_all = torch.empty(per * world, dtype=torch.uint8, device=dev) if (exchange == "peer" and rank == 0) else None
    prb_all = torch.empty(per * world, dtype=torch.float16, device=dev) if (exchange == "peer" and rank == 0) else None
    tick = torch.zeros(1, dtype=torch.int32, device=dev)
    eng.set_stream(stream.cuda_stream)
    if exchange == "peer":
        eng.close_peers()  # set_volume_device re-creates nothing (same size), but remap to be safe
    eng.set_volume_device(vol_dev.data_ptr(), (size, size, size))
    if exchange == "nccl":
        eng.bind_keys(keys.data_ptr())
    else:
        handles = [None] * world
        dist.all_gather_object(handles, eng.keys_ipc_handle())
        eng.open_peers(handles, rank)

    def one():
        with torch.cuda.stream(stream):
            vol_dev.copy_(vol_host.view(-1), non_blocking=True)
            if exchange == "nccl":
                keys.zero_()
            else:
                eng.reset()
            for it in items:
                eng.predict_range(it.d, it.s0, it.s1)
            if exchange == "nccl":
                dist.all_reduce(keys, op=dist.ReduceOp.MAX)
                if rank == 0:
                    eng.unpack_device(labels_dev.data_ptr(), probs_dev.data_ptr())
                    labels_h.copy_(labels_dev, non_blocking=True)
                    probs_h.copy_(probs_dev, non_blocking=True)
            else:
                dist.all_reduce(tick)
                eng.reduce_unpack_shard(v0, v1, labels_dev.data_ptr(), probs_dev.data_ptr())
                dist.all_reduce(tick)
                dist.gather(labels_dev, list(lab_all.split(per)) if rank == 0 else None, dst=0)
                dist.gather(probs_dev, list(prb_all.split(per)) if rank == 0 else None, dst=0)
                if rank == 0:
                    labels_h.copy_(lab_all[:nvox], non_blocking=True)
                    probs_h.copy_(prb_all[:nvox], non_blocking=True)
        torch.cuda.synchronize()

    one()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dist.barrier()
    t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if exchange == "peer":
        eng.close_peers()
    return {"value": nvox / float(t.item()), "unit": "voxels/s", "h2d_bytes_per_step": nvox * world,
            "d2h_bytes_per_step": 3 * nvox,
            "api": f"Engine.predict_range per rank + {exchange} exchange, host ndarray in/out"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="edge of the cubic synthetic volume (BASELINE: 1024)")
    ap.add_argument("--batch", type=int, default=0, help="slices per launch (0 = engine default)")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with CUDA events")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU key exchange: fused NVLink peer reduce+unpack (default) or NCCL all-reduce")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-oracle baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
