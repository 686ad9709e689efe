#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 pose hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

N > 1 is launched by torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.

Workloads (BASELINE.json configs):
  preproc_b256   configs[1]: PoseAugmentor + heat-map / regression head + composite loss kernels,
                 batch 256 of 256x256 RGB-D per GPU.  One step = one pass of that chain over one batch.
  cnn_train      configs[2]: CNNPoseEstimation full training step (forward, composite loss, backward, AdamW), bf16
                 tensor cores, batch 128 per GPU, data parallel (gradient all-reduce over NCCL overlapped with backward).
  vit_train      configs[3]: TransformerPoseEstimation full training step, batch 64 per GPU, data parallel.
The default line (preproc_b256) also carries `train` (both training steps at this N, so the 1/2/4/8-GPU runs record
their scaling) and `cnn_infer` (configs[4]: eval-mode CNN forward, batch sweep).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 42
B, H, W, J, HS, SIGMA = 256, 256, 256, 17, 256, 10.0
HEAD_IN, HEAD_HIDDEN = 1024, (1024, 512)


# ----------------------------------------------------------------------------------------------------
# synthetic inputs: seeded, identical for every implementation
# ----------------------------------------------------------------------------------------------------
def make_inputs(rank: int, batch: int = B):
    rng = np.random.default_rng(SEED + rank)
    img = rng.random((batch, 3, H, W), dtype=np.float32)
    dep = rng.random((batch, 1, H, W), dtype=np.float32)
    kp = rng.uniform(0.05, 0.95, (batch, J, 2)).astype(np.float32)
    kp[rng.random((batch, J)) < 0.05] = -1.0
    joints = rng.normal(0, 300, (batch, J, 3)).astype(np.float32)
    joints[:, :, 2] += 4000
    cam = np.tile(np.array([1145.0 * W / 1000, 1144.0 * H / 1000, W / 2.0, H / 2.0]), (batch, 1))
    feat = rng.normal(0, 1, (batch, HEAD_IN)).astype(np.float32)
    return dict(image=img, depth=dep, kp=kp, joints=joints, cam=cam, feat=feat)


def draw_params(aug, batch):
    np.random.seed(SEED)
    return aug.draw_params(batch)


# ----------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores (bounded sample)
# ----------------------------------------------------------------------------------------------------
def cpu_chain_samples_per_s(n_samples: int, threads: int, inputs=None, params=None):
    """Times the oracle port (C, one sample per call, ctypes releases the GIL) of the same chain:
    augment -> heat-map -> (head: fp32 numpy GEMMs) -> loss, on `n_samples` samples with `threads` threads."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    oracle.build()
    oracle.lib()
    inp = inputs or make_inputs(0, n_samples)
    pose = importlib.import_module("3dhumanposeestimation_b200")
    aug = pose.PoseAugmentor()
    if params is None:
        params = draw_params(aug, n_samples)
    rng = np.random.default_rng(SEED)
    dims = (HEAD_IN,) + HEAD_HIDDEN + (J * 3,)
    Ws = [rng.normal(0, 0.02, (dims[i + 1], dims[i])).astype(np.float32) for i in range(len(dims) - 1)]

    def one(i):
        o = oracle.augment_sample(inp["image"][i], inp["depth"][i], inp["kp"][i], inp["joints"][i], inp["cam"][i],
                                  params[i], aug.flags)
        oracle.heatmap(o["keypoints_2d"][None], HS, SIGMA)
        return o["joints_3d"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        gts = list(ex.map(one, range(n_samples)))
    x = inp["feat"][:n_samples]
    for k, w in enumerate(Ws):
        x = x @ w.T
        if k < len(Ws) - 1:
            x = x / (1.0 + np.exp(-x))
    oracle.pose_loss(x.reshape(n_samples, J, 3), np.stack(gts))
    dt = time.perf_counter() - t0
    return n_samples / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure Python and
    cannot travel to the GPU box, so this arm times the oracle port (oracle/pose_oracle.c, pinned bit-exact
    against the live reference) with all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 64
    inp = make_inputs(0, sample)
    for _ in range(args.warmup):
        cpu_chain_samples_per_s(min(sample, 16), cores, inp)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, _dt = cpu_chain_samples_per_s(sample, cores, inp)
        vals.append(v)
    ms = (time.perf_counter() - t0) * 1e3 / max(1, args.steps)
    value = float(np.median(vals))
    line = {
        "impl": "reference", "metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": {"workload": "preproc_b256", "note": f"each step = {sample} of the 256 samples of one batch"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} samples/step of the B=256 augment+heatmap+head+loss chain"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torchrun (see module docstring)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    importlib.import_module("3dhumanposeestimation_b200.build").build()
    pose = importlib.import_module("3dhumanposeestimation_b200")
    from importlib import import_module
    ops = import_module("3dhumanposeestimation_b200.ops")

    dutil = import_module("3dhumanposeestimation_b200.dist")
    inp = make_inputs(rank)          # weak scaling: every rank owns its own B samples, nothing is exchanged
    aug = pose.PoseAugmentor()
    params = draw_params(aug, B)
    crit = pose.ComprehensivePoseLoss()
    head = pose.PoseRegressionHead(HEAD_IN, J, hidden_dims=list(HEAD_HIDDEN), dropout=0.2, activation="silu").to(dev).eval()
    hm_gen = pose.GaussianHeatmapGenerator(J, HS, SIGMA).to(dev)

    # host (pinned) and device copies of one batch
    host = {k: torch.from_numpy(v).pin_memory() for k, v in inp.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    PAD = (308, 308)  # int(256 * 1.2) = 307 rows, width rounded up to a multiple of 4
    stream = torch.cuda.current_stream()

    def step(src):
        """One pass of the hot path over one batch; returns the 5 loss scalars (device tensor)."""
        a = aug.augment_batch(src["image"], src["depth"], src["kp"], src["joints"], src["cam"], params=params, pad_to=PAD)
        hm = hm_gen(a["keypoints_2d"])
        with torch.no_grad():
            pred = head(src["feat"])
        out5, grad = pose.loss.pose_loss_fwd_bwd(pred, a["joints_3d"], crit._weights())
        return out5, hm, a, grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_step = 3 + 1 + ops.MLP_HEAD_LAUNCHES(len(HEAD_HIDDEN) + 1) + 1

    # ---- device-resident timing: the step is captured once into a CUDA graph (10 kernels + 1 small memcpy node)
    #      and replayed K times, so the launch-bound Python/ctypes host path is outside the timed region ----------
    for _ in range(max(args.warmup, 3)):
        step(d)
    barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out5, hm, a, grad = step(d)
    for _ in range(max(args.warmup, 3)):
        graph.replay()
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        t_start.record()
        for i in range(args.steps):
            graph.replay()
        t_end.record()
        barrier()
    ms_total = t_start.elapsed_time(t_end)
    if aug.kernel_error_flag() != 0:
        raise RuntimeError("augment kernel reported a launch-geometry error")
    value = dutil.job_throughput(B * args.steps, ms_total, dev)   # all ranks' samples / slowest rank's device time
    ms_total = dutil.max_over_ranks(ms_total, dev)
    ms_per_step = ms_total / args.steps

    # ---- per-kernel device times for the roofline: each kernel group launched back to back R times between two
    #      CUDA events on its stream (a single bracketed launch would include host launch gaps) ------------------
    def timed(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    kp_aug, gt_aug = a["keypoints_2d"].clone(), a["joints_3d"].clone()
    pred0 = gt_aug + 25.0
    with torch.no_grad():
        kern_ms = {
            "augment": timed(lambda: aug.augment_batch(d["image"], d["depth"], d["kp"], d["joints"], d["cam"], params=params, pad_to=PAD)),
            "heatmap": timed(lambda: hm_gen(kp_aug)),
            "head": timed(lambda: head(d["feat"])),
            "loss": timed(lambda: pose.loss.pose_loss_fwd_bwd(pred0, gt_aug, crit._weights())),
        }

    # ---- end to end: pinned host inputs -> device, result scalars back, every step ------------------
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("image", "depth", "kp", "joints", "cam", "feat"))
    res_host = torch.empty(5, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    bufs = [{k: torch.empty_like(d[k]) for k in host} for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_steps(n):
        for f in free:
            f.record(stream)
        for i in range(n + 1):
            if i < n:  # stage batch i on the copy stream (overlaps the compute of batch i-1)
                s = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[s])
                    for k in host:
                        bufs[s][k].copy_(host[k], non_blocking=True)
                    ready[s].record(copy_stream)
            if i > 0:
                s = (i - 1) & 1
                stream.wait_event(ready[s])
                o5, *_ = step(bufs[s])
                res_host.copy_(o5, non_blocking=True)
                free[s].record(stream)
        torch.cuda.synchronize()

    e2e_steps(max(args.warmup, 3))
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_steps(args.steps)
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    e2e_value = dutil.job_throughput(B * args.steps, e2e_ms, dev)

    cnn = measure_cnn_infer(pose, dev, rank) if args.cnn else None
    vit_inf = measure_vit_infer(pose, dev, rank) if args.cnn else None
    train_res = None
    if args.train:
        train_res = {k: measure_train(pose, dev, rank, world, k, max(3, min(args.steps, 10)), 3) for k in ("cnn", "vit")}

    # ---- the same end-to-end loop fed with uint8 host pixels (4x fewer PCIe bytes; augment_batch's uint8 input is
    #      defined to reproduce the reference's fp32 sample p/255 bit for bit, tests/test_gpu_parity.py) -------------
    host8 = dict(host)
    host8["image"] = (host["image"] * 255.0).to(torch.uint8).pin_memory()
    host8["depth"] = (host["depth"] * 255.0).to(torch.uint8).pin_memory()
    h2d_u8 = sum(host8[k].numel() * host8[k].element_size() for k in ("image", "depth", "kp", "joints", "cam", "feat"))
    bufs8 = [{k: torch.empty_like(host8[k], device=dev) for k in host8} for _ in range(2)]

    def e2e_u8_steps(n):
        for f in free:
            f.record(stream)
        for i in range(n + 1):
            if i < n:
                s = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[s])
                    for k in host8:
                        bufs8[s][k].copy_(host8[k], non_blocking=True)
                    ready[s].record(copy_stream)
            if i > 0:
                s = (i - 1) & 1
                stream.wait_event(ready[s])
                o5, *_ = step(bufs8[s])
                res_host.copy_(o5, non_blocking=True)
                free[s].record(stream)
        torch.cuda.synchronize()

    e2e_u8_steps(3)
    barrier()
    t0 = time.perf_counter()
    e2e_u8_steps(args.steps)
    barrier()
    e2e_u8_value = dutil.job_throughput(B * args.steps, (time.perf_counter() - t0) * 1e3, dev)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        # dominant kernel group by time: the fused augmentation (pack + tables + cluster kernel).  Algorithmic bytes
        # (SURVEY.md 8d): 16 B/px fp32 in + 16 B per OUTPUT px + key-points/joints; padding writes are not counted.
        sizes = a["sizes"].cpu().numpy().astype(np.int64)
        aug_bytes = int(B * (16 * H * W) + 16 * int((sizes[:, 0] * sizes[:, 1]).sum()) + B * J * 20 * 2)
        achieved = aug_bytes / (kern_ms["augment"] * 1e-3) / 1e9
        hm_bytes = B * (J * HS * HS * 4 + J * 8)
        cpu_n = 256
        cores = os.cpu_count() or 1
        cpu_chain_samples_per_s(32, cores, inp)  # warm-up (page-in, thread pool)
        reps, cpu_dt = 0, 0.0
        while cpu_dt < 4.0 and reps < 64:  # bounded: a few seconds of wall clock on all host threads
            _v, dt1 = cpu_chain_samples_per_s(cpu_n, cores, inp, params)
            cpu_dt += dt1
            reps += 1
        cpu_v = cpu_n * reps / cpu_dt
        line = {
            "metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8+f32 (head: bf16 x bf16 -> f32)", "data": "synthetic",
            "config": {"workload": "preproc_b256", "batch_per_gpu": B, "image": [H, W], "heatmap": [HS, SIGMA],
                       "chain": "PoseAugmentor(all stages) -> GaussianHeatmap(256, sigma 10) -> PoseRegressionHead "
                                "1024-1024-512-51 -> ComprehensivePoseLoss fwd+bwd",
                       "l2": "inputs (268 MB/batch) and outputs (1.5 GB/batch) exceed the 126 MB L2",
                       "launch": "device-resident value: CUDA-graph replay of the 10-kernel step; e2e: eager launches"},
            "clocks": clocks.summary(),
            "gpu_launches": launches_per_step * args.steps,
            "kernel_ms": kern_ms,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 20,
                    "note": "fp32 pinned host batch (reference sample schema) -> device every step, double-buffered",
                    "uint8_input": {"value": e2e_u8_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d_u8),
                                    "note": "same call with uint8 decoded pixels (bit-identical outputs), 4x fewer PCIe bytes"}},
            "roofline": {"kernel": "pose_augment_batch: aug_pack + aug_tables + aug_fused_kernel<5> (dominant by time)",
                         "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": 67.8e6 + 331.2e6 + 335e6, "traffic_note": "ncu dram bytes r+w per launch: fused 399 MB + pack 335 MB "
                         "(profiles/r01_*.csv); algorithmic bytes per launch = bytes below",
                         "bytes_per_launch": aug_bytes, "peak_source": peak_src,
                         "others": {"heatmap_planes_kernel<float>": {"bytes": hm_bytes, "GB/s": hm_bytes / (kern_ms["heatmap"] * 1e-3) / 1e9,
                                                                   "frac": hm_bytes / (kern_ms["heatmap"] * 1e-3) / 1e9 / hbm_peak},
                                    "loss": {"bytes": B * 632, "ms": kern_ms["loss"], "note": "latency bound at B=256"},
                                    "head (3 tcgen05 GEMMs + cast)": {"ms": kern_ms["head"]}}},
            "cnn_infer": cnn,
            "vit_infer": vit_inf,
            "train": train_res,
            "cpu_baseline": {"value": cpu_v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{reps} x the same B=256 batch ({cpu_dt:.1f} s wall, {cpu_dt * cores:.0f} core-s), oracle port (C) on all host threads"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# training step (configs[2], configs[3]): forward + loss + backward + AdamW, data parallel
# ----------------------------------------------------------------------------------------------------
TRAIN_CFG = {
    "cnn": dict(batch=128, gflop_per_sample=49.68, params=26_920_792),
    "vit": dict(batch=64, gflop_per_sample=212.2, params=147_774_515),
}


def build_train_model(pose, kind, dev):
    import torch
    torch.manual_seed(SEED)
    if kind == "cnn":
        cfg = pose.ModelConfig("cnn", image_size=(H, W), heatmap_size=HS)     # reference defaults incl. dropout 0.2
        return pose.CNNPoseEstimation(cfg).to(dev).train(), cfg
    # reference defaults (dropout 0.1 / attention dropout 0.1 / head dropout 0.25); random init instead of timm weights
    cfg = pose.ModelConfig("transformer", image_size=(H, W), vit_pretrained=False)
    return pose.TransformerPoseEstimation(cfg).to(dev).train(), cfg


def measure_train(pose, dev, rank, world, kind, steps, warmup, cpu_baseline=True):
    """One rank's share of the data-parallel training step; returns the whole-job numbers (max over ranks)."""
    import torch
    import torch.distributed as dist
    from importlib import import_module
    train = import_module("3dhumanposeestimation_b200.train")
    dutil = import_module("3dhumanposeestimation_b200.dist")
    spec = TRAIN_CFG[kind]
    Bn = spec["batch"]
    model, cfg = build_train_model(pose, kind, dev)
    if world > 1:
        train.broadcast_parameters(model)
    g = torch.Generator().manual_seed(SEED + rank)
    host = dict(image=torch.rand(Bn, 3, H, W, generator=g).pin_memory(), depth=torch.rand(Bn, 1, H, W, generator=g).pin_memory(),
                kp=(torch.rand(Bn, J, 2, generator=g) * 0.9 + 0.05).pin_memory(),
                gt=(torch.randn(Bn, J, 3, generator=g) * 300).pin_memory())
    d = {k: v.to(dev) for k, v in host.items()}
    tr = train.Trainer(model, pose.ComprehensivePoseLoss(), lr=1e-3, weight_decay=0.01)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warmup, 3)):
        o5 = tr.step(d["image"], d["depth"], d["kp"], d["gt"])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clocks:
        e0.record()
        for _ in range(steps):
            o5 = tr.step(d["image"], d["depth"], d["kp"], d["gt"])
        e1.record()
        barrier()
    ms = dutil.max_over_ranks(e0.elapsed_time(e1), dev)
    value = world * Bn * steps / ms * 1e3
    plan = model.plan(Bn, dev)
    launches = plan.launches + 3          # forward + backward kernels of the last step, loss, AdamW, workspace clear
    # end to end: the batch comes from pinned host memory every step, the 5 loss scalars go back
    res_host = torch.empty(5, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    stream = torch.cuda.current_stream()
    bufs = [{k: torch.empty_like(d[k]) for k in host} for _ in range(2)]
    ready, free = [torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_steps(n):
        for f in free:
            f.record(stream)
        for i in range(n + 1):
            if i < n:
                sidx = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[sidx])
                    for k in host:
                        bufs[sidx][k].copy_(host[k], non_blocking=True)
                    ready[sidx].record(copy_stream)
            if i > 0:
                sidx = (i - 1) & 1
                stream.wait_event(ready[sidx])
                o = tr.step(bufs[sidx]["image"], bufs[sidx]["depth"], bufs[sidx]["kp"], bufs[sidx]["gt"])
                res_host.copy_(o, non_blocking=True)
                free[sidx].record(stream)
        torch.cuda.synchronize()

    e2e_steps(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(steps)
    barrier()
    e2e_ms = dutil.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf = value * spec["gflop_per_sample"] / 1e3
    out = {"model": kind, "batch_per_gpu": Bn, "n_gpus": world, "samples_per_s": value, "ms_per_step": ms / steps,
           "tflops": tf, "tensor_frac_of_measured_sustained": tf / world / peaks.get("bf16_tflops_sustained", 1344.7),
           "gflop_per_sample_train": spec["gflop_per_sample"], "launches_per_step": launches,
           "loss_total": float(o5[4].item()), "clocks": clocks.summary(),
           "e2e": {"value": world * Bn * steps / e2e_ms * 1e3, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": 20},
           "dtype": "bf16 activations / weights on the tensor cores, fp32 accumulate, fp32 master weights + AdamW state",
           "data": "synthetic", "parallelism": f"dp{world}" if world > 1 else "single GPU",
           "optimizer": "fused AdamW lr 1e-3 wd 0.01 (main.py:154-156), accumulation_steps 1"}
    out["note"] = "reference default configuration incl. dropout (CNN head 0.2; ViT 0.1 / attention 0.1 / head 0.25)"
    if rank == 0 and cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_train_baseline(pose, kind, cfg)
    del tr, model, plan, d, bufs
    torch.cuda.empty_cache()
    return out


def cpu_train_baseline(pose, kind, cfg):
    """The reference's CPU path of the same training step (fp32 PyTorch on the host cores): the oracle restatement, pinned
    to the live reference by tests/golden (the reference itself cannot travel to the GPU box)."""
    import torch
    try:
        from oracle import torch_models as tm
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        bs = 8 if kind == "cnn" else 2
        model = (pose.CNNPoseEstimation(cfg) if kind == "cnn" else pose.TransformerPoseEstimation(cfg))
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        names = [n for n, _ in model.named_parameters()]
        g = torch.Generator().manual_seed(SEED)
        img, dep = torch.rand(bs, 3, H, W, generator=g), torch.rand(bs, 1, H, W, generator=g)
        kp, gt = torch.rand(bs, J, 2, generator=g) * 0.9 + 0.05, torch.randn(bs, J, 3, generator=g) * 300
        iu = torch.triu_indices(J, J, 1)
        pd = lambda t: torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]   # noqa: E731

        def step():
            sdg = {k: (v.clone().requires_grad_() if k in names else v) for k, v in sd.items()}
            po = tm.cnn_forward(sdg, cfg, img, dep, kp, train=True) if kind == "cnn" else tm.vit_forward(sdg, cfg, img, dep, kp)
            d = po - gt
            ((d ** 2).mean() + d.abs().mean() + 100.0 * (pd(po) - pd(gt)).abs().mean() + d[:, 0].abs().mean()).backward()
        step()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            step()
        dt = (time.perf_counter() - t0) / n
        return {"value": bs / dt, "unit": "samples/s", "cores": cores, "kind": "port",
                "sample": f"fp32 forward + loss + backward, B={bs}, {n} repetitions ({dt * n:.1f} s); optimizer step not included"}
    except Exception as exc:
        return {"error": repr(exc)}


def run_train(args, kind):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    importlib.import_module("3dhumanposeestimation_b200.build").build()
    pose = importlib.import_module("3dhumanposeestimation_b200")
    r = measure_train(pose, dev, rank, world, kind, args.steps, args.warmup)
    if rank == 0:
        line = {"metric": "train samples/sec", "value": r["samples_per_s"], "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{kind}_train", "batch_per_gpu": r["batch_per_gpu"], "image": [H, W],
                           "parallelism": r["parallelism"], "l2": "activations of one step (GBs) exceed the 126 MB L2"},
                "clocks": r["clocks"], "gpu_launches": r["launches_per_step"] * args.steps, "e2e": r["e2e"],
                "roofline": {"bound": "tensor", "achieved": r["tflops"] / world, "peak": 1344.7, "unit": "TFLOP/s",
                             "frac": r["tensor_frac_of_measured_sustained"], "traffic": None,
                             "note": "whole step: nominal training FLOPs (SURVEY.md 8d) / step time, per GPU"},
                "cpu_baseline": r.get("cpu_baseline"), "train": r}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_cnn_infer(pose, dev, rank, batches=(1, 8, 32, 128, 256, 512), reps=10):
    """Eval-mode CNNPoseEstimation forward at 256x256 (BASELINE configs[4]): samples/s per batch size, CUDA events."""
    import torch
    cfg = pose.ModelConfig("cnn", image_size=(H, W), heatmap_size=HS)
    torch.manual_seed(SEED)
    model = pose.CNNPoseEstimation(cfg).to(dev).eval()
    out = {"config": "CNNPoseEstimation eval forward, bf16 tensor cores (fp32 accumulate), 256x256, random init, synthetic",
           "gflop_per_sample": 16.559, "samples_per_s": {}, "ms": {}}
    g = torch.Generator(device="cpu").manual_seed(SEED + rank)
    for bs in batches:
        img = torch.rand(bs, 3, H, W, generator=g).to(dev)
        dep = torch.rand(bs, 1, H, W, generator=g).to(dev)
        kp = (torch.rand(bs, J, 2, generator=g) * 0.9 + 0.05).to(dev)
        with torch.no_grad():
            for _ in range(3):
                model(img, dep, kp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                model(img, dep, kp)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out["ms"][str(bs)] = ms
        out["samples_per_s"][str(bs)] = bs / ms * 1e3
        plan = model._plans.pop((bs, dev.index), None)
        out["launches_per_forward"] = plan.launches + 1 if plan is not None else None
        del plan, img, dep
        torch.cuda.empty_cache()
    best = max(out["samples_per_s"].values())
    out["tflops_at_best"] = best * 16.559e9 / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    out["tensor_frac_of_measured_sustained"] = out["tflops_at_best"] / peaks.get("bf16_tflops_sustained", 1400.0)
    if rank == 0:
        # the reference's CPU path for the same forward (configs[0] flavour): fp32 oracle restatement on the host cores
        try:
            from oracle import torch_models as tm
            torch.set_num_threads(os.cpu_count() or 1)
            sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
            bs = 8
            img, dep, kp = torch.rand(bs, 3, H, W), torch.rand(bs, 1, H, W), torch.rand(bs, J, 2) * 0.9 + 0.05
            with torch.no_grad():
                tm.cnn_forward(sd, cfg, img, dep, kp)
                t0 = time.perf_counter()
                n = 3
                for _ in range(n):
                    tm.cnn_forward(sd, cfg, img, dep, kp)
                dt = (time.perf_counter() - t0) / n
            out["cpu_baseline"] = {"value": bs / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"fp32 eval forward, B={bs}, {n} repetitions ({dt * n:.1f} s)"}
        except Exception as exc:  # the checker is optional for the measurement itself
            out["cpu_baseline"] = {"error": repr(exc)}
    return out


def measure_vit_infer(pose, dev, rank, batches=(1, 8, 64, 256), reps=5):
    """Eval-mode TransformerPoseEstimation forward at 256x256 (row H for the ViT): samples/s per batch size, CUDA events."""
    import torch
    cfg = pose.ModelConfig("transformer", image_size=(H, W), vit_pretrained=False)
    torch.manual_seed(SEED)
    model = pose.TransformerPoseEstimation(cfg).to(dev).eval()
    out = {"config": "TransformerPoseEstimation eval forward, bf16 tensor cores (fp32 accumulate), 256x256, random init, synthetic",
           "gflop_per_sample": 70.73, "samples_per_s": {}, "ms": {}}
    g = torch.Generator(device="cpu").manual_seed(SEED + rank)
    for bs in batches:
        img = torch.rand(bs, 3, H, W, generator=g).to(dev)
        dep = torch.rand(bs, 1, H, W, generator=g).to(dev)
        kp = (torch.rand(bs, J, 2, generator=g) * 0.9 + 0.05).to(dev)
        with torch.no_grad():
            for _ in range(2):
                model(img, dep, kp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                model(img, dep, kp)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out["ms"][str(bs)] = ms
        out["samples_per_s"][str(bs)] = bs / ms * 1e3
        model._plans.clear()
        torch.cuda.empty_cache()
    out["tflops_at_best"] = max(out["samples_per_s"].values()) * 70.73e9 / 1e12
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="preproc_b256")
    ap.add_argument("--no-cnn", dest="cnn", action="store_false", help="skip the CNN inference sweep")
    ap.add_argument("--no-train", dest="train", action="store_false", help="skip the training-step measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("cnn_train", "vit_train"):
        run_train(args, args.workload.split("_")[0])
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
