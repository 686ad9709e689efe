#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 pose hot path (contract: DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

N > 1 is launched by torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.

The metric (BASELINE.json) is **train samples/sec (CNN, ViT)**; the headline line is therefore

  cnn_train      configs[2]: CNNPoseEstimation full training step (forward, composite loss, backward, AdamW), bf16 tensor
                 cores, batch 128 per GPU, 256x256, data parallel (NCCL gradient all-reduce overlapped with backward);

and the same line carries, as complete second records with their own clocks / roofline / e2e / baselines,

  vit_train      configs[3]: TransformerPoseEstimation full training step, batch 64 per GPU, data parallel;
  preproc_b256   configs[1]: PoseAugmentor + heat-map / regression head + composite loss kernels, batch 256 (N = 1 only);
  cnn_infer / vit_infer   configs[4]: eval-mode forward batch sweep 1..1024 (N = 1 only).

Comparators, all measured in the same run on the same box:
  gpu_eager_baseline  the same training step as plain PyTorch on the B200: oracle/torch_models functional forward under
                      torch.autocast(bfloat16), channels-last (CNN) / fused SDPA attention (ViT), torch loss, autograd,
                      torch.optim.AdamW(fused=True) -- i.e. the library path (cuDNN / cuBLASLt / flash attention);
  cpu_baseline        the reference's CPU path of the step (fp32 PyTorch on all host cores, bounded sample);
  --impl reference    the driver's reference arm: the same CPU training step, one fixed thread pool, on the same `config`.
`--workload vit_train | preproc_b256` makes that workload the headline instead.
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

TRAIN_CFG = {
    "cnn": dict(batch=128, gflop_per_sample=49.68, params=26_920_792, cpu_micro=16, cpu_micros=2),
    "vit": dict(batch=64, gflop_per_sample=212.2, params=147_774_515, cpu_micro=4, cpu_micros=2),
}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm": p["hbm_gbs"], "tc": p["bf16_tflops_sustained"], "tc_burst": p.get("bf16_tflops"), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tc": 1400.0, "tc_burst": 1590.0, "src": "fallback"}


def train_config(kind, world):
    """The `config` object of a training workload: identical for the B200 arm and the reference arm."""
    spec = TRAIN_CFG[kind]
    return {"workload": f"{kind}_train",
            "model": "CNNPoseEstimation" if kind == "cnn" else "TransformerPoseEstimation (ViT-B/16 backbone)",
            "batch_per_gpu": spec["batch"], "global_batch": spec["batch"] * world, "image": [H, W],
            "parallelism": f"dp{world}", "gradient_exchange": "none (1 GPU)" if world == 1 else
            "all-reduce of the flat gradient in " + os.environ.get("POSE_GRAD_WIRE", "bf16") + " buckets of <= 64 MB, overlapped with backward",
            "optimizer": "AdamW lr 1e-3 wd 0.01 (main.py:154-156), accumulation_steps 1",
            "step": "forward + ComprehensivePoseLoss + backward + AdamW (src/train.py:76-119), reference default "
                    "dropout rates, random init, synthetic 256x256 RGB-D batches",
            "l2": "activations of one step (GBs) exceed the 126 MB L2; inputs are re-read from HBM every step"}


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


def train_batch(torch, rank, n):
    g = torch.Generator().manual_seed(SEED + rank)
    return dict(image=torch.rand(n, 3, H, W, generator=g), depth=torch.rand(n, 1, H, W, generator=g),
                kp=torch.rand(n, J, 2, generator=g) * 0.9 + 0.05, gt=torch.randn(n, J, 3, generator=g) * 300)


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
# the training step as plain PyTorch (oracle restatement): CPU reference arm and GPU eager comparator
# ----------------------------------------------------------------------------------------------------
class TorchStep:
    """forward + composite loss + backward + AdamW of one model in plain PyTorch over oracle/torch_models (the
    restatement pinned to the live reference by tests/golden).  Used ONLY as a comparator: fp32 on the host cores
    (cpu_baseline / --impl reference) or autocast-bf16 on the B200 (gpu_eager_baseline).  Never on the product path."""

    def __init__(self, pose, kind, device, autocast=False):
        import torch
        from oracle import torch_models as tm
        self.torch, self.tm, self.kind, self.dev, self.autocast = torch, tm, kind, device, autocast
        torch.manual_seed(SEED)
        if kind == "cnn":
            self.cfg = pose.ModelConfig("cnn", image_size=(H, W), heatmap_size=HS)
            model = pose.CNNPoseEstimation(self.cfg)
        else:
            self.cfg = pose.ModelConfig("transformer", image_size=(H, W), vit_pretrained=False)
            model = pose.TransformerPoseEstimation(self.cfg)
        names = {n for n, _ in model.named_parameters()}
        self.sd = {}
        for k, v in model.state_dict().items():
            t = v.detach().clone().to(device)
            if autocast and kind == "cnn" and t.dim() == 4:
                t = t.contiguous(memory_format=torch.channels_last)
            if k in names:
                t.requires_grad_(True)
            self.sd[k] = t
        del model
        self.params = [t for k, t in self.sd.items() if k in names]
        self.opt = torch.optim.AdamW(self.params, lr=1e-3, weight_decay=0.01, fused=bool(autocast))

    def micro(self, b, scale=1.0):
        torch, tm = self.torch, self.tm
        img, dep = b["image"], b["depth"]
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if self.autocast else _Null()
        with ctx:
            if self.kind == "cnn":
                if self.autocast:
                    img = img.contiguous(memory_format=torch.channels_last)
                pred, stats = tm.cnn_forward(self.sd, self.cfg, img, dep, b["kp"], train=True, return_stats=True,
                                             dropout=float(self.cfg.regression_dropout))
                for k, v in stats.items():          # nn.BatchNorm2d updates its running buffers in place
                    self.sd[k] = v
            else:
                pred = tm.vit_forward(self.sd, self.cfg, img, dep, b["kp"], train=True, sdpa=True)
        loss = tm.composite_loss(pred.float(), b["gt"]) * scale
        loss.backward()
        return loss

    def step(self, micros):
        for b in micros:
            loss = self.micro(b, 1.0 / len(micros))
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def cpu_train_step_rate(pose, kind, steps, warmup):
    """The reference's CPU path of the training step: fp32, every host core in ONE fixed thread pool.  A step here is a
    bounded sample of the workload's batch (micro-batches with gradient accumulation, then one AdamW step)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = TRAIN_CFG[kind]
    mb, nm = spec["cpu_micro"], spec["cpu_micros"]
    ts = TorchStep(pose, kind, torch.device("cpu"))
    micros = [train_batch(torch, 100 + i, mb) for i in range(nm)]
    for _ in range(warmup):
        ts.step(micros[:1])
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ts.step(micros)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": mb * nm / med, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{mb * nm} of the {spec['batch']} samples of one batch per step, as {nm} micro-batches of {mb} with "
                      f"gradient accumulation + one AdamW step; fp32 PyTorch restatement of the reference step "
                      f"(oracle/torch_models, pinned by tests/golden), median of {steps} steps, {med:.2f} s/step",
            "s_per_step": med}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the training step (rank 0 only).  The reference is
    pure Python + PyTorch and cannot travel to the GPU box, so the arm times the oracle restatement of the same step
    (kind "port") on all host cores, on this arm's `config`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pose = importlib.import_module("3dhumanposeestimation_b200")
    kind = args.workload.split("_")[0] if args.workload.endswith("_train") else "cnn"
    t0 = time.perf_counter()
    r = cpu_train_step_rate(pose, kind, args.steps, max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": "train samples/sec", "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": train_config(kind, args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.extras and kind == "cnn" and time.perf_counter() - t0 < 120:
        v = cpu_train_step_rate(pose, "vit", max(2, min(args.steps, 5)), 1)
        line["vit_train"] = {"impl": "reference", "metric": "train samples/sec", "value": v["value"], "unit": "samples/s",
                             "config": train_config("vit", args.gpus),
                             "cpu_baseline": {k: v[k] for k in ("value", "unit", "cores", "kind", "sample")}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# training step (configs[2], configs[3]): forward + loss + backward + AdamW, data parallel
# ----------------------------------------------------------------------------------------------------
def build_train_model(pose, kind, dev):
    import torch
    torch.manual_seed(SEED)
    if kind == "cnn":
        cfg = pose.ModelConfig("cnn", image_size=(H, W), heatmap_size=HS)     # reference defaults incl. dropout 0.2
        return pose.CNNPoseEstimation(cfg).to(dev).train(), cfg
    # reference defaults (dropout 0.1 / attention dropout 0.1 / head dropout 0.25); random init instead of timm weights
    cfg = pose.ModelConfig("transformer", image_size=(H, W), vit_pretrained=False)
    return pose.TransformerPoseEstimation(cfg).to(dev).train(), cfg


def step_roofline(pose, tr, plan, d, spec, world, value, pk):
    """One extra (untimed) step with every launch bracketed by CUDA events: per-family device time, achieved TFLOP/s and
    GB/s against the measured peaks.  The top-level fraction is the whole step's nominal training FLOPs over step time."""
    import torch
    trace_mod = importlib.import_module("3dhumanposeestimation_b200.trace")
    graph_mode, tr.use_graph = tr.use_graph, False      # the traced step is launched kernel by kernel (events around each)
    with trace_mod.PlanTrace(plan) as t:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.step(d["image"], d["depth"], d["kp"], d["gt"])
        e1.record()
        fam, ent = t.summary()
    tr.use_graph = graph_mode
    traced_ms = e0.elapsed_time(e1)
    kern_ms = sum(v["ms"] for v in fam.values())
    groups = {}
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        gth = {"ms": round(v["ms"], 3), "share_of_kernel_time": round(v["ms"] / kern_ms, 4), "launches": v["launches"]}
        if v["flops"] > 0:
            gth["tflops"] = v["flops"] / (v["ms"] * 1e-3) / 1e12
            gth["frac_of_tensor_peak"] = gth["tflops"] / pk["tc"]
        if v["bytes"] > 0:
            gth["gbs"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
            gth["frac_of_hbm_peak"] = gth["gbs"] / pk["hbm"]
            gth["bytes"] = v["bytes"]
        groups[k] = gth
    top = sorted(ent.items(), key=lambda kv: -kv[1]["ms"])[:12]
    hbm_groups = {k: g for k, g in groups.items() if k in ("batchnorm", "depthwise", "layernorm", "adamw")}
    dom = max(hbm_groups.items(), key=lambda kv: kv[1]["ms"]) if hbm_groups else None
    tf = value * spec["gflop_per_sample"] / 1e3 / world
    out = {"kernel": "whole training step (nominal training FLOPs of SURVEY.md 8d / step time, per GPU)",
           "bound": "tensor", "achieved": tf, "peak": pk["tc"], "unit": "TFLOP/s", "frac": tf / pk["tc"], "traffic": None,
           "peak_source": pk["src"] + " sustained cuBLAS bf16 (a kernel timed inside a long step)",
           "timing": "one extra step with CUDA events around every launch on the launching stream (in this run, no profiler)",
           "traced_step_ms": traced_ms, "sum_of_kernel_ms": kern_ms, "groups": groups,
           "top_entry_points_ms": {k: round(v["ms"], 3) for k, v in top}}
    if dom is not None:
        out["dominant_hbm_group"] = {"group": dom[0], "bound": "hbm", "achieved": dom[1]["gbs"], "peak": pk["hbm"],
                                     "unit": "GB/s", "frac": dom[1]["frac_of_hbm_peak"], "bytes_per_step": dom[1]["bytes"],
                                     "ms_per_step": dom[1]["ms"]}
    return out


def gpu_eager_step_rate(pose, kind, dev, d, steps=5, warmup=3):
    """The same step as plain PyTorch on this B200 (see module docstring).  Returns samples/s and ms/step."""
    import torch
    spec = TRAIN_CFG[kind]
    try:
        ts = TorchStep(pose, kind, dev, autocast=True)
        for _ in range(warmup):
            ts.step([d])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = ts.step([d])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"value": spec["batch"] / ms * 1e3, "unit": "samples/s", "ms_per_step": ms, "loss": float(loss.item()),
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
               "what": "oracle/torch_models functional forward under torch.autocast(bfloat16) + torch composite loss + "
                       "autograd + torch.optim.AdamW(fused=True); " +
                       ("channels-last cuDNN convolutions" if kind == "cnn" else "F.scaled_dot_product_attention (fused) "
                        "with the reference's dropout sites") + f"; torch {torch.__version__}, same B={spec['batch']}"}
        del ts
        torch.cuda.empty_cache()
        return out
    except Exception as exc:  # the comparator must never take the measurement down
        torch.cuda.empty_cache()
        return {"error": repr(exc)[:300]}


def measure_train(pose, dev, rank, world, kind, steps, warmup, extras=True):
    """One rank's share of the data-parallel training step; returns a complete record (whole-job numbers, max over ranks)."""
    import torch
    import torch.distributed as dist
    from importlib import import_module
    train = import_module("3dhumanposeestimation_b200.train")
    dutil = import_module("3dhumanposeestimation_b200.dist")
    spec = TRAIN_CFG[kind]
    Bn = spec["batch"]
    pk = peaks()
    model, cfg = build_train_model(pose, kind, dev)
    if world > 1:
        train.broadcast_parameters(model)
    host = {k: v.pin_memory() for k, v in train_batch(torch, rank, Bn).items()}
    d = {k: v.to(dev) for k, v in host.items()}
    tr = train.Trainer(model, pose.ComprehensivePoseLoss(), lr=1e-3, weight_decay=0.01)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(warmup, 3)
    for _ in range(warmup):
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
    launches = tr.launches_per_step(plan)
    # ---- end to end: the batch comes from pinned host memory every step, the 5 loss scalars go back ----------------
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
    roof = step_roofline(pose, tr, plan, d, spec, world, value, pk)
    rec = {"metric": "train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps,
           "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16", "data": "synthetic", "config": train_config(kind, world), "clocks": clocks.summary(),
           "gpu_launches": launches * steps, "launches_per_step": launches, "launch_mode": tr.launch_mode(),
           "e2e": {"value": world * Bn * steps / e2e_ms * 1e3, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": 20,
                   "note": "fp32 pinned host batch -> device every step (double-buffered copy stream), Trainer.step, 5 loss "
                           "scalars back to pinned host memory; wall clock, max over ranks"},
           "roofline": roof, "loss_total": float(o5[4].item()),
           "precision": "bf16 activations / weights on the tensor cores, fp32 accumulate, fp32 master weights + AdamW state"}
    del bufs
    if rank == 0 and world == 1 and extras:
        tr_mem = torch.cuda.max_memory_allocated() / 2 ** 30
        rec["peak_mem_gb"] = tr_mem
        # free the product's buffers before the comparator allocates its own activations
        del tr, plan
        model._plans.clear()
        del model
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        rec["gpu_eager_baseline"] = gpu_eager_step_rate(pose, kind, dev, d)
        if "value" in rec["gpu_eager_baseline"]:
            rec["vs_gpu_eager"] = value / rec["gpu_eager_baseline"]["value"]
        try:
            c = cpu_train_step_rate(pose, kind, 3, 1)
            rec["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:
            rec["cpu_baseline"] = {"error": repr(exc)[:300]}
    else:
        del tr, plan
        model._plans.clear()
        del model
    del d
    torch.cuda.empty_cache()
    return rec


# ----------------------------------------------------------------------------------------------------
# configs[1]: augment + heat-map + head + loss chain at batch 256
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


def measure_preproc(pose, dev, rank, world, steps, warmup, cpu=True):
    import torch
    import torch.distributed as dist
    from importlib import import_module
    ops = import_module("3dhumanposeestimation_b200.ops")
    dutil = import_module("3dhumanposeestimation_b200.dist")
    pk = peaks()
    inp = make_inputs(rank)          # weak scaling: every rank owns its own B samples, nothing is exchanged
    aug = pose.PoseAugmentor()
    params = draw_params(aug, B)
    crit = pose.ComprehensivePoseLoss()
    head = pose.PoseRegressionHead(HEAD_IN, J, hidden_dims=list(HEAD_HIDDEN), dropout=0.2, activation="silu").to(dev).eval()
    hm_gen = pose.GaussianHeatmapGenerator(J, HS, SIGMA).to(dev)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in inp.items()}
    d = {k: v.to(dev) for k, v in host.items()}
    PAD = (308, 308)  # int(256 * 1.2) = 307 rows, width rounded up to a multiple of 4
    stream = torch.cuda.current_stream()

    def step(src):
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

    launches_per_step = aug.launches_per_batch() + 1 + ops.MLP_HEAD_LAUNCHES(len(HEAD_HIDDEN) + 1) + 1
    warmup = max(warmup, 3)
    for _ in range(warmup):
        step(d)
    barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out5, hm, a, grad = step(d)
    for _ in range(warmup):
        graph.replay()
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clocks:
        t_start.record()
        for i in range(steps):
            graph.replay()
        t_end.record()
        barrier()
    ms_total = t_start.elapsed_time(t_end)
    if aug.kernel_error_flag() != 0:
        raise RuntimeError("augment kernel reported a launch-geometry error")
    value = dutil.job_throughput(B * steps, ms_total, dev)
    ms_per_step = dutil.max_over_ranks(ms_total, dev) / steps

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
    res_host = torch.empty(5, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_rate(hostd):
        bufs = [{k: torch.empty_like(hostd[k], device=dev) for k in hostd} for _ in range(2)]

        def run(n):
            for f in free:
                f.record(stream)
            for i in range(n + 1):
                if i < n:  # stage batch i on the copy stream (overlaps the compute of batch i-1)
                    s = i & 1
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(free[s])
                        for k in hostd:
                            bufs[s][k].copy_(hostd[k], non_blocking=True)
                        ready[s].record(copy_stream)
                if i > 0:
                    s = (i - 1) & 1
                    stream.wait_event(ready[s])
                    o5, *_ = step(bufs[s])
                    res_host.copy_(o5, non_blocking=True)
                    free[s].record(stream)
            torch.cuda.synchronize()
        run(3)
        barrier()
        t0 = time.perf_counter()
        run(steps)
        barrier()
        nbytes = sum(hostd[k].numel() * hostd[k].element_size() for k in hostd)
        return dutil.job_throughput(B * steps, (time.perf_counter() - t0) * 1e3, dev), int(nbytes)

    e2e_value, h2d = e2e_rate(host)
    host8 = dict(host)
    host8["image"] = (host["image"] * 255.0).to(torch.uint8).pin_memory()
    host8["depth"] = (host["depth"] * 255.0).to(torch.uint8).pin_memory()
    e2e_u8_value, h2d_u8 = e2e_rate(host8)

    sizes = a["sizes"].cpu().numpy().astype(np.int64)
    aug_bytes = int(B * (16 * H * W) + 16 * int((sizes[:, 0] * sizes[:, 1]).sum()) + B * J * 20 * 2)
    achieved = aug_bytes / (kern_ms["augment"] * 1e-3) / 1e9
    hm_bytes = B * (J * HS * HS * 4 + J * 8)
    rec = {"metric": "samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps,
           "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u8+f32 (head: bf16 x bf16 -> f32)", "data": "synthetic",
           "config": {"workload": "preproc_b256", "batch_per_gpu": B, "image": [H, W], "heatmap": [HS, SIGMA],
                      "chain": "PoseAugmentor(all stages) -> GaussianHeatmap(256, sigma 10) -> PoseRegressionHead "
                               "1024-1024-512-51 -> ComprehensivePoseLoss fwd+bwd",
                      "l2": "inputs (268 MB/batch) and outputs (1.5 GB/batch) exceed the 126 MB L2",
                      "launch": "device-resident value: CUDA-graph replay of the step; e2e: eager launches"},
           "clocks": clocks.summary(), "gpu_launches": launches_per_step * steps, "kernel_ms": kern_ms,
           "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 20,
                   "note": "fp32 pinned host batch (reference sample schema) -> device every step, double-buffered",
                   "uint8_input": {"value": e2e_u8_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_u8,
                                   "note": "same call with uint8 decoded pixels (bit-identical outputs), 4x fewer PCIe bytes"}},
           "roofline": {"kernel": "pose_augment_batch (dominant by time)", "bound": "hbm", "achieved": achieved,
                        "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"], "traffic": None,
                        "bytes_per_launch": aug_bytes, "peak_source": pk["src"],
                        "others": {"heatmap_planes_kernel<float>": {"bytes": hm_bytes, "GB/s": hm_bytes / (kern_ms["heatmap"] * 1e-3) / 1e9,
                                                                  "frac": hm_bytes / (kern_ms["heatmap"] * 1e-3) / 1e9 / pk["hbm"]},
                                   "loss": {"bytes": B * 632, "ms": kern_ms["loss"], "note": "latency bound at B=256"},
                                   "head (3 tcgen05 GEMMs + cast)": {"ms": kern_ms["head"]}}}}
    if rank == 0 and cpu:
        cores = os.cpu_count() or 1
        cpu_chain_samples_per_s(32, cores, inp)  # warm-up (page-in, thread pool)
        reps, cpu_dt = 0, 0.0
        while cpu_dt < 4.0 and reps < 64:  # bounded: a few seconds of wall clock on all host threads
            _v, dt1 = cpu_chain_samples_per_s(B, cores, inp, params)
            cpu_dt += dt1
            reps += 1
        rec["cpu_baseline"] = {"value": B * reps / cpu_dt, "unit": "samples/s", "cores": cores, "kind": "port",
                               "sample": f"{reps} x the same B=256 batch ({cpu_dt:.1f} s wall), oracle port (C) on all host threads"}
    del graph
    torch.cuda.empty_cache()
    return rec


# ----------------------------------------------------------------------------------------------------
# configs[4]: eval-mode forward, batch sweep 1..1024
# ----------------------------------------------------------------------------------------------------
def measure_infer(pose, dev, rank, kind, batches, reps=5):
    """Eval-mode forward at 256x256 (BASELINE configs[4] / row H): samples/s and latency per batch size, CUDA events.
    Small batches are also replayed from a CUDA graph (the launch-bound regime)."""
    import torch
    torch.manual_seed(SEED)
    if kind == "cnn":
        cfg = pose.ModelConfig("cnn", image_size=(H, W), heatmap_size=HS)
        model = pose.CNNPoseEstimation(cfg).to(dev).eval()
        gf = 16.559
    else:
        cfg = pose.ModelConfig("transformer", image_size=(H, W), vit_pretrained=False)
        model = pose.TransformerPoseEstimation(cfg).to(dev).eval()
        gf = 70.73
    out = {"config": f"{type(model).__name__} eval forward, bf16 tensor cores (fp32 accumulate), 256x256, random init, synthetic",
           "gflop_per_sample": gf, "samples_per_s": {}, "ms": {}, "graph_ms": {}}
    g = torch.Generator(device="cpu").manual_seed(SEED + rank)
    for bs in batches:
        try:
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
                if bs <= 8:
                    try:
                        gr = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gr):
                            model(img, dep, kp)
                        gr.replay()
                        torch.cuda.synchronize()
                        e0.record()
                        for _ in range(20):
                            gr.replay()
                        e1.record()
                        torch.cuda.synchronize()
                        out["graph_ms"][str(bs)] = e0.elapsed_time(e1) / 20
                        del gr
                    except Exception as exc:
                        out["graph_ms"][str(bs)] = "error: " + repr(exc)[:120]
            out["ms"][str(bs)] = ms
            best = min(ms, out["graph_ms"].get(str(bs), ms) if isinstance(out["graph_ms"].get(str(bs), ms), float) else ms)
            out["samples_per_s"][str(bs)] = bs / best * 1e3
        except Exception as exc:
            out["ms"][str(bs)] = "error: " + repr(exc)[:120]
        model._plans.clear()
        img = dep = kp = None
        torch.cuda.empty_cache()
    vals = [v for v in out["samples_per_s"].values() if isinstance(v, float)]
    if vals:
        out["tflops_at_best"] = max(vals) * gf * 1e9 / 1e12
        out["tensor_frac_of_measured_sustained"] = out["tflops_at_best"] / peaks()["tc"]
    del model
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch N > 1 with torchrun (see module docstring)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    importlib.import_module("3dhumanposeestimation_b200.build").build()
    pose = importlib.import_module("3dhumanposeestimation_b200")
    extras = args.extras
    t_start = time.perf_counter()
    if args.workload == "preproc_b256":
        line = measure_preproc(pose, dev, rank, world, args.steps, args.warmup)
    else:
        kind = args.workload.split("_")[0]
        line = measure_train(pose, dev, rank, world, kind, args.steps, args.warmup, extras)
        other = "vit" if kind == "cnn" else "cnn"
        if extras:
            line[f"{other}_train"] = measure_train(pose, dev, rank, world, other, args.steps, args.warmup, extras)
        if extras and world == 1:
            line["preproc_b256"] = measure_preproc(pose, dev, rank, world, args.steps, args.warmup)
            line["cnn_infer"] = measure_infer(pose, dev, rank, "cnn", (1, 2, 4, 8, 32, 128, 512, 1024))
            line["vit_infer"] = measure_infer(pose, dev, rank, "vit", (1, 2, 4, 8, 64, 256, 1024))
    line["bench_wall_s"] = time.perf_counter() - t_start
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cnn_train", choices=["cnn_train", "vit_train", "preproc_b256"])
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="headline workload only (no second records, comparators or sweeps)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
