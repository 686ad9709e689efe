"""Per-launch tracing of a launch plan (NVTX ranges and kernel-level timing are new tooling; the reference
only has tqdm / TensorBoard scalars, src/train.py:42-65).

``PlanTrace`` wraps the ``call`` method of a launch plan (``CnnTrainPlan``, ``CnnInferencePlan``, ``VitPlan``): every C-ABI
entry point the plan launches is bracketed by two CUDA events on the launching stream and, optionally, by an NVTX range
named after the entry point (visible in nsys / ncu ``--nvtx`` captures).  ``work()`` attaches the algorithmic FLOPs or
bytes of a launch (from the shapes in its argument list), so that a whole step can be summarised per kernel family
against the tensor-core and HBM rooflines without a profiler attached.
"""
from __future__ import annotations

import torch


def _conv_out(n, k, s, d, p):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


def work(name, a):
    """(family, flops, bytes) of one launch of entry point `name` with argument tuple `a` (stream excluded)."""
    if name in ("pose_gemm_bf16_ex",):
        M, N, K = a[4], a[5], a[6]
        return "tensor", 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)
    if name == "pose_gemm_bf16_tr":
        M, N, K = a[6], a[7], a[8]
        return "tensor", 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)
    if name == "pose_gemm_bf16":
        M, N, K = a[7], a[8], a[9]
        return "tensor", 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)
    if name in ("pose_conv2d_bf16", "pose_conv2d_wgrad_bf16"):
        o = 1 if name == "pose_conv2d_wgrad_bf16" else 0
        B, H, W, cin = a[1 + o], a[2 + o], a[3 + o], a[4 + o]
        co, kh, kw, st, dil, pad = (a[6:12] if o == 0 else a[6:12])
        Ho, Wo = _conv_out(H, kh, st, dil, pad), _conv_out(W, kw, st, dil, pad)
        return "tensor", 2.0 * B * Ho * Wo * co * kh * kw * cin, 2.0 * (B * H * W * cin + B * Ho * Wo * co)
    if name == "pose_attention_bf16":
        B, h, Nq, Nk, hd = a[4:9]
        return "attention", 4.0 * B * h * Nq * Nk * hd, 2.0 * B * h * hd * (2 * Nq + 2 * Nk)
    if name == "pose_attention_bwd_bf16":
        B, h, Nq, Nk, hd = a[10:15]
        return "attention", 10.0 * B * h * Nq * Nk * hd, 2.0 * B * h * hd * (4 * Nq + 4 * Nk)
    if name == "pose_bn_stats_bf16":
        return "batchnorm", 0.0, 2.0 * a[1] * a[2]
    if name == "pose_bn_apply_bf16":
        return "batchnorm", 0.0, (4.0 + (2.0 if a[6] else 0.0)) * a[1] * a[2]
    if name == "pose_bn_bwd_bf16":
        return "batchnorm", 0.0, 10.0 * a[3] * a[4]
    if name == "pose_bn_bwd_from_dz_bf16":               # dz and y read, dY written (the reduction ran in the producer)
        return "batchnorm", 0.0, 6.0 * a[3] * a[4]
    if name == "pose_gate_bwd_apply_bn_bf16":            # dOut and the BatchNorm's y read, dz written: the reduction pass of
        return "batchnorm", 0.0, 6.0 * a[4] * a[5] * a[6]    # that BatchNorm is this launch (counted with its family)
    if name == "pose_dwconv3x3_bnbwd_bf16":
        B, H, W, C = a[1:5]
        return "depthwise", 0.0, 2.0 * B * C * 3 * H * W
    if name in ("pose_bn_finalize", "pose_bn_finalize_parts"):
        return "batchnorm", 0.0, 0.0
    if name in ("pose_dwconv3x3_bf16", "pose_dwconv3x3_bn_stats_bf16"):
        B, H, W, C = a[1:5]
        st = a[7] if name == "pose_dwconv3x3_bf16" else a[6]
        return "depthwise", 0.0, 2.0 * B * C * (H * W + ((H - 1) // st + 1) * ((W - 1) // st + 1))
    if name == "pose_dwconv3x3_bwd_bf16":
        B, H, W, C, st = a[3:8]
        return "depthwise", 0.0, 2.0 * B * C * (2 * H * W + ((H - 1) // st + 1) * ((W - 1) // st + 1))
    if name == "pose_adamw_step":
        return "adamw", 0.0, 30.0 * a[5]
    if name in ("pose_layernorm_bf16", "pose_layernorm_bwd_bf16"):
        return "layernorm", 0.0, (4.0 if name == "pose_layernorm_bf16" else 8.0) * a[4] * a[10]
    return "other", 0.0, 0.0


class PlanTrace:
    """with PlanTrace(plan, nvtx=False) as t: ...run a step...;  t.summary() -> per-family totals."""

    def __init__(self, *plans, nvtx: bool = False, timing: bool = True):
        self.plans, self.nvtx, self.timing = [p for p in plans if p is not None], nvtx, timing
        self.log = []
        self._orig = []

    def __enter__(self):
        trace = self

        def make(orig):
            def traced(plan_self, name, *args):
                if trace.nvtx:
                    torch.cuda.nvtx.range_push(name)
                if trace.timing:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    orig(plan_self, name, *args)
                    e1.record()
                    trace.log.append((name, args, e0, e1))
                else:
                    orig(plan_self, name, *args)
                if trace.nvtx:
                    torch.cuda.nvtx.range_pop()
            return traced
        for cls in {type(p) for p in self.plans}:
            self._orig.append((cls, cls.call))
            cls.call = make(cls.call)
        return self

    def __exit__(self, *exc):
        for cls, orig in self._orig:
            cls.call = orig
        self._orig = []

    def record_extra(self, name, args, fn):
        """Times a launch that does not go through plan.call (loss, optimizer)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.log.append((name, args, e0, e1))
        return out

    def launches(self):
        """[(entry point, family, ms, flops, bytes)] in launch order (synchronises)."""
        torch.cuda.synchronize()
        out = []
        for name, args, e0, e1 in self.log:
            fam, fl, by = work(name, args)
            out.append((name, fam, e0.elapsed_time(e1), fl, by))
        return out

    def summary(self):
        """{family: {"ms", "launches", "flops", "bytes"}} plus per-entry-point times."""
        fam, ent = {}, {}
        for name, f, ms, fl, by in self.launches():
            d = fam.setdefault(f, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
            d["ms"] += ms
            d["launches"] += 1
            d["flops"] += fl
            d["bytes"] += by
            e = ent.setdefault(name, {"ms": 0.0, "launches": 0})
            e["ms"] += ms
            e["launches"] += 1
        return fam, ent
