#!/usr/bin/env python
"""Benchmark of the prediction hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1..cfg5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[2] ("cfg3"), the configuration the metric is quoted on:
12-direction ('high' quality) U-Net/ResNet-34, 4 classes, synthetic 1024^3 uint8 volume (fits one B200).
--config selects the other BASELINE configurations (architecture, classes, quality, shape):
  cfg1 U-Net/R34 low C=2 256^3 | cfg2 U-Net/R34 medium C=4 512^3 | cfg4 U-Net++/ResNeXt-50 high C=6 512^3 |
  cfg5 DeepLabV3+/R50 medium C=4 (2048,2048,512).
One "step" = one full prediction of the volume.  The same volume is sharded over N GPUs (strong
scaling): work items = (direction, slice range), one exchange of the packed keys.

  value : volume already resident in HBM -> label + fp16 probability volumes in HBM (CUDA events on
          the engine's stream, max over ranks).
  e2e   : host ndarray -> host ndarrays, copies inside the timed region.  N = 1: the reference-facing call
          itself, VolSeg2dPredictor(model file)._predict_*(ndarray).  N > 1 (one process per GPU): every
          rank uploads 1/N of the volume from pinned memory, the parts are all-gathered over NVLink, and
          every rank downloads ITS shard of labels + probabilities into one shared page-locked result.
  roofline : the tcgen05 convolution class (achieved = algorithmic conv FLOPs / summed CUDA-event
          durations of its launches); `roofline_classes` adds the HBM-bound classes (slicer, stem + pool,
          head + merge) the same way.
  cpu_baseline : the CPU oracle (restatement of the reference path) on a bounded sample, extrapolated
          (cfg1: timed in full); a reported baseline, not a target.
  result_sha256 : SHA-256 of the label + probability volumes of the e2e leg (identical at every N).

--impl reference runs ONLY the CPU oracle port (the reference itself cannot be imported in this image:
h5py / segmentation_models_pytorch / albumentations are absent) with all host threads.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# BASELINE.json `configs`, in order (cfg3 = the one the metric is quoted on)
CONFIGS = {
    "cfg1": dict(arch="U_NET", oracle_arch="unet", encoder="resnet34", classes=2, quality="low", shape=(256, 256, 256),
                 name="U-Net/ResNet-34, low quality (1 direction), 2 classes, synthetic 256^3 uint8 volume"),
    "cfg2": dict(arch="U_NET", oracle_arch="unet", encoder="resnet34", classes=4, quality="medium", shape=(512, 512, 512),
                 name="U-Net/ResNet-34, medium quality (3 directions), 4 classes, synthetic 512^3 uint8 volume"),
    "cfg3": dict(arch="U_NET", oracle_arch="unet", encoder="resnet34", classes=4, quality="high", shape=(1024, 1024, 1024),
                 name="U-Net/ResNet-34, high quality (12 directions), 4 classes, synthetic 1024^3 uint8 volume"),
    "cfg4": dict(arch="U_NET_PLUS_PLUS", oracle_arch="unetplusplus", encoder="resnext50_32x4d", classes=6, quality="high",
                 shape=(512, 512, 512),
                 name="U-Net++/ResNeXt-50_32x4d, high quality (12 directions), 6 classes, synthetic 512^3 uint8 volume"),
    "cfg5": dict(arch="DEEPLABV3_PLUS", oracle_arch="deeplabv3plus", encoder="resnet50", classes=4, quality="medium",
                 shape=(2048, 2048, 512),
                 name="DeepLabV3+/ResNet-50, medium quality (3 directions), 4 classes, synthetic (2048,2048,512) uint8 volume"),
}
QUALITY_MASK = {"low": 0b001, "medium": 0b111, "high": (1 << 12) - 1}
QUALITY_DIRS = {"low": 1, "medium": 3, "high": 12}
METRIC_CFG3 = "voxels/sec, 12-direction U-Net prediction of a 1024^3 volume"


def env_int(name, default):
    return int(os.environ.get(name, default))


def resolve_config(args):
    cfg = dict(CONFIGS[args.config])
    if args.size:  # development aid: a smaller cube of the same configuration
        cfg["shape"] = (args.size,) * 3
        cfg["name"] = cfg["name"].split(", synthetic")[0] + f", synthetic {args.size}^3 uint8 volume"
    cfg["key"] = args.config
    cfg["dir_mask"] = QUALITY_MASK[cfg["quality"]]
    cfg["n_dirs"] = QUALITY_DIRS[cfg["quality"]]
    if args.config == "cfg3" and not args.size:
        cfg["metric"] = METRIC_CFG3
    else:
        z, y, x = cfg["shape"]
        cfg["metric"] = f"voxels/sec, {cfg['n_dirs']}-direction {cfg['name'].split(',')[0]} prediction of a {z}x{y}x{x} volume"
    return cfg


def config_dict(cfg):
    """Identical for both arms (the driver compares them textually)."""
    return {"workload": cfg["name"], "baseline_config": cfg["key"],
            "weights": "random init of the named architecture (seed 0), BN statistics randomised"}


def synth_volume(shape):
    # integers(0, 256) uint8 volume, fixed seed
    return np.random.default_rng(20240).integers(0, 256, size=shape, dtype=np.uint8)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json: sustained bf16, HBM copy)"
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


def make_models(cfg, want_oracle=False):
    """Random init of the named architecture with randomised BN statistics."""
    import torch

    from volume_segmantics_b200.plan import B200SegmentationModel

    torch.manual_seed(0)
    model = B200SegmentationModel(cfg["arch"], cfg["encoder"], cfg["classes"])
    g = torch.Generator().manual_seed(1)
    sd = model.state_dict()
    for key, (shape, (role, _)) in model.spec.param_shapes().items():
        if role in ("bn_w", "bn_var"):
            sd[key].copy_(torch.rand(shape, generator=g) + 0.5)
        elif role in ("bn_b", "bn_mean"):
            sd[key].copy_(torch.randn(shape, generator=g) * 0.1)
    return model


# ----------------------------------------------------------------------------- CPU oracle leg
def cpu_oracle_sample(cfg, threads, per_axis):
    """Time the oracle (a) in full for cfg1, else (b) on `per_axis` slices of every AXIS orientation
    the configuration slices along (strided views included, as the reference gathers them) + one merge
    sample, extrapolated linearly.  Returns (voxels/s, sample description)."""
    import torch

    from oracle import predict_oracle as po
    from oracle.smp_models import make_random_model

    torch.set_num_threads(threads)
    model = make_random_model(cfg["oracle_arch"], cfg["encoder"], cfg["classes"], seed=0)
    pred = po.OraclePredictor(model, cfg["classes"], batch_size=4)
    Z, Y, X = cfg["shape"]
    nvox = Z * Y * X
    if cfg["key"] == "cfg1" and nvox <= 256 ** 3:
        vol = synth_volume(cfg["shape"])
        t0 = time.perf_counter()
        pred.predict_single_axis(vol, True, po.AXIS_Z)
        dt = time.perf_counter() - t0
        return nvox / dt, f"the whole {Z}x{Y}x{X} volume, single axis, batch 4 (timed in full: {dt:.1f} s)"
    rng = np.random.default_rng(1)
    axes = (0,) if cfg["quality"] == "low" else (0, 1, 2)
    total, parts = 0.0, []
    for a in axes:
        # a slab thick enough for `per_axis` slices along axis a, full size in the slice plane
        shp = [Z, Y, X]
        shp[a] = per_axis
        slab = rng.integers(0, 256, size=shp, dtype=np.uint8)
        t0 = time.perf_counter()
        pred.predict_single_axis(slab, True, a)
        t_slice = (time.perf_counter() - t0) / per_axis
        n_slices = (Z, Y, X)[a] * (cfg["n_dirs"] // len(axes))  # rot90 variants slice the same planes (transposed)
        total += n_slices * t_slice
        parts.append(f"axis {'ZYX'[a]} {t_slice:.3f} s/slice")
    n_merges = cfg["n_dirs"] - 1
    t_merge_vox = 0.0
    if n_merges:
        mz = max(1, min(Z, (32 << 20) // (Y * X)))
        pc = rng.random((2, mz, Y, X)).astype(np.float16)
        lc = np.zeros((2, mz, Y, X), np.uint8)
        t0 = time.perf_counter()
        po.merge_vols_in_mem(pc, lc)
        t_merge_vox = (time.perf_counter() - t0) / (mz * Y * X)
        total += n_merges * nvox * t_merge_vox
    sample = (f"{per_axis} slices per axis orientation ({', '.join(parts)}) through the fp32 CPU oracle (batch 4) + one "
              f"fp16 merge sample ({t_merge_vox * 1e9:.1f} ns/voxel), extrapolated linearly to {cfg['n_dirs']} directions "
              f"+ {n_merges} merges")
    return nvox / total, sample


def run_reference(args, cfg):
    if env_int("RANK", 0) != 0:
        return
    threads = os.cpu_count() or 1
    vals, sample = [], ""
    warm, steps = args.warmup, args.steps
    if cfg["key"] == "cfg1":  # timed in full (tens of seconds per step): bound the run
        warm, steps = min(warm, 1), min(steps, 3)
    for i in range(warm + steps):
        v, sample = cpu_oracle_sample(cfg, threads, 2)
        if i >= warm:
            vals.append(v)
    value = float(np.mean(vals))
    nvox = int(np.prod(cfg["shape"]))
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * nvox / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": config_dict(cfg),
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU leg
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from volume_segmantics_b200 import _lib, sharding
    from volume_segmantics_b200.engine import get_engine
    from volume_segmantics_b200.plan import conv_macs_per_pixel

    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    shape = tuple(cfg["shape"])
    nvox = int(np.prod(shape))

    model = make_models(cfg)
    vol_host = torch.from_numpy(synth_volume(shape)).pin_memory()
    eng = get_engine(local)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    eng.load_model(model)
    if args.batch:
        eng.set_batch(args.batch)
    vol_dev = vol_host.to(dev)
    torch.cuda.synchronize()
    eng.set_volume_device(vol_dev.data_ptr(), shape)
    dirs = sharding.direction_list(cfg["dir_mask"], skip_duplicates=True)
    items = sharding.partition(shape, dirs, world, granule=8)[rank]

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
    per = shards[0][1] - shards[0][0]
    out_n = nvox if exchange != "peer" else per
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
    # kernel launch bracketed by CUDA events, for the rooflines of the kernel classes.
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
        lt = torch.tensor([launches], dtype=torch.float64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)  # gpu_launches is the whole job's count, like `value`
        launches = int(lt.item())
    ms_step = float(t.item()) / args.steps
    value = nvox / (ms_step * 1e-3)

    # ---- rooflines, this rank's launches -------------------------------------------------------
    peak_tf, peak_gbs, peak_src = measured_peaks()
    stem_macs = 49 * 64 / 4.0  # the 7x7 stem is timed in its own (HBM-bound) class
    macs_px = conv_macs_per_pixel(model.spec) - stem_macs
    padded_px = sum(it.cost for it in items)
    C = cfg["classes"]
    conv_ms, conv_n = stages["conv_tc"]
    conv_flops = 2.0 * macs_px * padded_px * args.steps
    roofline = None
    traffic = None  # DRAM bytes per conv launch from the committed ncu capture of one batch (profiles/)
    tpath = ROOT / "profiles" / "r02_conv_traffic.json"
    if tpath.exists() and cfg["key"] == "cfg3" and not args.size:
        traffic = json.loads(tpath.read_text()).get("dram_bytes_per_conv_launch")
    if conv_ms > 0:
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "tcgen05 conv kernels (conv_halo_kernel, conv_halo2_kernel<>, conv_tc_kernel<>)", "achieved": ach,
                    "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, average over the conv launches of one "
                                      "batch of 128 slices of 1024^2, the launch configuration of this run (profiles/r02_ncu_batch128_dram.csv)" if traffic else None,
                    "peak_source": peak_src,
                    "launches": conv_n, "avg_launch_ms": conv_ms / max(1, conv_n),
                    "flops_per_launch": conv_flops / max(1, conv_n),
                    "share_of_step": conv_ms / ms_prof,
                    "measured": "second pass of the same K steps with per-launch CUDA events "
                                f"({ms_prof / args.steps:.1f} ms/step with events vs {ms / args.steps:.1f} without)",
                    "other_stage_ms_per_step": {k: v[0] / args.steps for k, v in stages.items() if k != "conv_tc"}}
    # HBM-bound classes: algorithmic bytes per padded pixel / event time
    head_factor = model.spec.layers[-1].factor
    per_px = {
        "slicer": (3.0, "1 B volume read + 2 B activation written"),
        "stem": (2.0 + 32.0 + 8.0, "stem 7x7/2 + fused 3x3/2 max-pool: 2 B in + 64 ch x 2 B / 4 out + 64 ch x 2 B / 16 pooled"),
        "head": (4.0 * C / (head_factor ** 2) + 16.0, f"{C} fp32 logits (at 1/{head_factor} resolution) + 8 B key read + 8 B key written"),
    }
    classes = {}
    for name, (bpp, what) in per_px.items():
        ms_c, n_c = stages[name]
        if name == "stem":
            ms_c += stages["pool"][0]
        if ms_c > 0:
            gbs = bpp * padded_px * args.steps / (ms_c * 1e-3) / 1e9
            classes[name] = {"bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
                             "bytes_per_padded_pixel": bpp, "what": what, "ms_per_step": ms_c / args.steps, "launches": n_c}

    # ---- e2e through the public API with host buffers ------------------------------
    e2e, digest = run_e2e(args, cfg, eng, model, vol_host, stream, world, rank, items, keys, dev, exchange, shards)

    # ---- CPU baseline: bounded sample on rank 0, N = 1 only -------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v, sample = cpu_oracle_sample(cfg, threads, 2)
        cpu = {"value": v, "unit": "voxels/s", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp16" if _lib.act_dtype() == torch.float16 else "bf16", "data": "synthetic",
            "config": config_dict(cfg),
            "details": {"directions_computed": len(dirs),
                        "note": "directions 3,6,9,10 duplicate 1,4,7,0 image-for-image and can never win the first-max merge "
                                "(SURVEY.md 3.3); they are skipped and NOT counted in the roofline FLOPs",
                        "l2": f"inputs larger than L2 ({nvox >> 20} MiB volume + {nvox >> 17} MiB keys per step)",
                        "accumulate": "fp32", "parallelism": f"slice-range sharding over {world} GPU(s)",
                        "exchange": {"none": "single GPU: unpack only",
                                     "peer": "fused max-reduce + unpack of each rank's voxel shard over NVLink peer memory (CUDA IPC); result sharded over ranks",
                                     "nccl": "ncclAllReduce(max) of the 8 B/voxel key volume, rank 0 unpacks"}[exchange]},
            "roofline": roofline, "roofline_classes": classes, "cpu_baseline": cpu, "e2e": e2e, "result_sha256": digest,
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sha256_arrays(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(memoryview(np.ascontiguousarray(a)).cast("B"))
    return h.hexdigest()


def run_e2e(args, cfg, eng, model, vol_host, stream, world, rank, items, keys, dev, exchange, shards):
    """Host ndarray in -> host ndarrays out, copies inside the timed region.  Returns (e2e dict, sha256)."""
    import torch
    import torch.distributed as dist

    shape = tuple(cfg["shape"])
    nvox = int(np.prod(shape))
    vol_np = vol_host.numpy()
    steps = max(1, args.steps)
    if world == 1:
        # exactly what a user of the reference does: a .pytorch file, VolSeg2dPredictor, one _predict_* call
        from types import SimpleNamespace

        import volume_segmantics.utilities.base_data_utils as utils
        from volume_segmantics.model.operations.vol_seg_2d_predictor import VolSeg2dPredictor

        struc = {"type": utils.ModelType[cfg["arch"]], "encoder_name": cfg["encoder"], "encoder_weights": None,
                 "in_channels": 1, "classes": cfg["classes"]}
        with tempfile.TemporaryDirectory() as tmp:
            path = Path(tmp) / "bench_model.pytorch"
            torch.save({"model_state_dict": model.state_dict(), "model_struc_dict": struc, "label_codes": {}}, path)
            eng.bind_keys(0)
            eng.set_stream(0)
            pred = VolSeg2dPredictor(str(path), SimpleNamespace(cuda_device=eng.device))
        call = {"low": lambda v: pred._predict_single_axis(v), "medium": pred._predict_3_ways_max_probs,
                "high": pred._predict_12_ways_max_probs}[cfg["quality"]]
        call(vol_np)  # warm-up (weights lowered for this module, allocations)
        labels = probs = None
        t0 = time.perf_counter()
        for _ in range(steps):
            labels = probs = None  # the caller is done with the previous result: its pinned block is reused
            labels, probs = call(vol_np)
        dt = (time.perf_counter() - t0) / steps
        assert labels.shape == vol_np.shape and probs.dtype == np.float16
        fn = {"low": "_predict_single_axis", "medium": "_predict_3_ways_max_probs", "high": "_predict_12_ways_max_probs"}[cfg["quality"]]
        return ({"value": nvox / dt, "unit": "voxels/s", "h2d_bytes_per_step": nvox, "d2h_bytes_per_step": 3 * nvox, "steps": steps,
                 "api": f"VolSeg2dPredictor(<model>.pytorch, settings).{fn}(ndarray) -> (uint8, float16) ndarrays; "
                        "model file loading and weight lowering happen once, before the timed calls"},
                sha256_arrays(labels, probs))

    # ---- N > 1: one process per GPU ---------------------------------------------------------
    per = shards[0][1] - shards[0][0]
    v0, v1 = shards[rank]
    # ONE page-locked result shared by all ranks (a /dev/shm mapping registered with the CUDA driver in
    # every process): each rank downloads its own shard over its own PCIe link
    tag = os.environ.get("MASTER_PORT", "0")
    lab_path, prb_path = f"/dev/shm/vsb200_bench_{tag}_labels", f"/dev/shm/vsb200_bench_{tag}_probs"
    if rank == 0:
        for p, n in ((lab_path, nvox), (prb_path, 2 * nvox)):
            with open(p, "wb") as f:
                f.truncate(n)
    dist.barrier()
    labels_h = torch.from_file(lab_path, shared=True, size=nvox, dtype=torch.uint8)
    probs_h = torch.from_file(prb_path, shared=True, size=nvox, dtype=torch.float16)
    rt = torch.cuda.cudart()
    for tsr in (labels_h, probs_h):
        rc = rt.cudaHostRegister(tsr.data_ptr(), tsr.numel() * tsr.element_size(), 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed: {rc}")
    vol_dev = torch.empty(per * world, dtype=torch.uint8, device=dev)  # all-gather output (>= nvox)
    vol_flat = vol_host.view(-1)
    n_out = nvox if exchange == "nccl" else per
    labels_dev = torch.empty(n_out, dtype=torch.uint8, device=dev)
    probs_dev = torch.empty(n_out, dtype=torch.float16, device=dev)
    tick = torch.zeros(1, dtype=torch.int32, device=dev)
    eng.set_stream(stream.cuda_stream)
    if exchange == "peer":
        eng.close_peers()
    eng.set_volume_device(vol_dev.data_ptr(), shape)
    if exchange == "nccl":
        eng.bind_keys(keys.data_ptr())
    else:
        handles = [None] * world
        dist.all_gather_object(handles, eng.keys_ipc_handle())
        eng.open_peers(handles, rank)
    my = vol_dev[rank * per:(rank + 1) * per]

    def one():
        with torch.cuda.stream(stream):
            if v1 > v0:
                my[: v1 - v0].copy_(vol_flat[v0:v1], non_blocking=True)  # 1/N of the volume over this rank's PCIe link
            dist.all_gather_into_tensor(vol_dev, my)  # the other parts over NVLink
            if exchange == "nccl":
                keys.zero_()
            else:
                eng.reset()
            for it in items:
                eng.predict_range(it.d, it.s0, it.s1)
            if exchange == "nccl":
                dist.all_reduce(keys, op=dist.ReduceOp.MAX)
                eng.unpack_device(labels_dev.data_ptr(), probs_dev.data_ptr())
                if v1 > v0:
                    labels_h[v0:v1].copy_(labels_dev[v0:v1], non_blocking=True)
                    probs_h[v0:v1].copy_(probs_dev[v0:v1], non_blocking=True)
            else:
                dist.all_reduce(tick)
                eng.reduce_unpack_shard(v0, v1, labels_dev.data_ptr(), probs_dev.data_ptr())
                dist.all_reduce(tick)
                if v1 > v0:
                    labels_h[v0:v1].copy_(labels_dev[: v1 - v0], non_blocking=True)
                    probs_h[v0:v1].copy_(probs_dev[: v1 - v0], non_blocking=True)
        torch.cuda.synchronize()

    one()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dist.barrier()  # every shard of the last result is in the shared host buffer
    t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    digest = None
    if rank == 0:
        digest = sha256_arrays(labels_h.numpy().reshape(shape), probs_h.numpy().reshape(shape))
    dist.barrier()
    for tsr in (labels_h, probs_h):
        rt.cudaHostUnregister(tsr.data_ptr())
    if exchange == "peer":
        eng.close_peers()
    dist.barrier()
    if rank == 0:
        for p in (lab_path, prb_path):
            try:
                os.unlink(p)
            except OSError:
                pass
    return ({"value": nvox / float(t.item()), "unit": "voxels/s", "h2d_bytes_per_step": nvox, "d2h_bytes_per_step": 3 * nvox,
             "steps": steps,
             "api": f"one process per GPU: 1/{world} of the volume uploaded per rank + NCCL all-gather over NVLink, "
                    f"Engine.predict_range per work item, {exchange} exchange, every rank downloads its shard of labels + "
                    "fp16 probabilities into one shared page-locked host result"},
            digest)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS), help="BASELINE.json configuration (default cfg3)")
    ap.add_argument("--size", type=int, default=0, help="development aid: edge of a smaller cubic volume of the same configuration")
    ap.add_argument("--batch", type=int, default=0, help="slices per launch (0 = engine default)")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with CUDA events")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU key exchange: fused NVLink peer reduce+unpack (default) or NCCL all-reduce")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-oracle baseline sample")
    args = ap.parse_args()
    cfg = resolve_config(args)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
