"""The per-batch training step on B200 (reference: src/train.py:76-119, optimizer main.py:154-156).

``Trainer.step`` is the body of the reference's batch loop -- forward, composite loss, backward, (every
``accumulation_steps`` batches) AdamW step + zero_grad -- with the reference's six ``.item()`` syncs replaced by
one 5-float device buffer.  It drives the model's launch plan directly (the same kernels ``loss.backward()``
reaches through autograd; the autograd route stays available for drop-in use).

Data parallelism: one process per GPU, every rank owns its own samples and a full replica; the only
exchange is the gradient all-reduce.  The flat fp32 gradient buffer is reduced in a few large buckets in the
order the backward pass completes them (model tail first), each bucket on a communication stream as soon as its
section of the backward has been enqueued, so NCCL over NVLink overlaps the remaining backward kernels.
BatchNorm statistics stay per replica (the reference has no SyncBN).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import _lib  # noqa: F401
from .dist import allreduce_range
from .loss import pose_loss_fwd_bwd
from .optim import AdamW


class Trainer:
    def __init__(self, model, criterion, optimizer=None, accumulation_steps=1, lr=1e-3, weight_decay=0.01,
                 process_group=None, bucket_bytes=64 << 20, grad_wire=None):
        self.model, self.crit = model, criterion
        self.opt = optimizer if optimizer is not None else AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
        self.accum = int(accumulation_steps)
        self.micro = 0
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.pg = process_group
        self.bucket_bytes = bucket_bytes
        self.comm_stream = torch.cuda.Stream() if self.world > 1 else None
        self._pending = []
        self.out5 = None
        # gradient exchange format: "bf16" halves the bytes on NVLink (each finished section of the flat fp32 gradient is
        # cast to a bf16 staging buffer on the communication stream, all-reduced there, and read by the fused AdamW
        # directly -- the same compression as torch DDP's bf16_compress_hook); "fp32" reduces the fp32 buffer in place
        if grad_wire is None:
            grad_wire = os.environ.get("POSE_GRAD_WIRE", "bf16")       # A/B switch for measurements
        if grad_wire not in ("bf16", "fp32"):
            raise ValueError("grad_wire must be 'bf16' or 'fp32'")
        self.grad_wire = grad_wire
        self._grad16 = None

    # ---- gradient exchange ---------------------------------------------------------------------------
    def _section_done(self, flat, lo, hi):
        """Gradients in flat[lo:hi) are final: all-reduce them on the communication stream (bucketed)."""
        if self.world == 1 or hi <= lo:
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            if self.grad_wire == "bf16":
                if self._grad16 is None or self._grad16.numel() != flat.numel:
                    self._grad16 = torch.zeros(flat.numel, dtype=torch.bfloat16, device=flat.grad.device)
                    self.opt.grad16 = self._grad16
                _lib.check(_lib.lib().pose_cast_f32_bf16(flat.grad.data_ptr() + 4 * lo, self._grad16.data_ptr() + 2 * lo,
                                                         hi - lo, self.comm_stream.cuda_stream), "pose_cast_f32_bf16")
                allreduce_range(self._grad16, lo, hi, self.bucket_bytes // 2, self.pg)
            else:
                allreduce_range(flat.grad, lo, hi, self.bucket_bytes // 4, self.pg)

    def _wait_comm(self):
        if self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    # ---- one batch -----------------------------------------------------------------------------------
    def step(self, images, depths, keypoints_2d, gt_joints):
        """Returns the device tensor [mse, l1, inter_joint, abs_root, total] of this batch (no host sync)."""
        m = self.model
        B = images.shape[0]
        plan = m.plan(B, images.device)
        last = (self.micro + 1) % self.accum == 0
        out = plan.forward(images, depths, keypoints_2d, save=True)
        pred = out.view(B, -1, 3)
        # loss forward + d(total / accum)/d(pred) in one kernel (src/train.py:86-92)
        out5, grad = pose_loss_fwd_bwd(pred, gt_joints, self.crit._weights(), want_grad=True, grad_scale=1.0 / self.accum)
        hook = self._section_done if (last and self.world > 1) else None
        plan.backward(grad.view(B, -1), section_done=hook)
        self.micro += 1
        if last:
            self._wait_comm()
            self.opt.grad_scale = 1.0 / self.world          # all-reduce(sum) / world = mean over replicas
            self.opt.step()                                  # also clears the flat gradient buffer
        self.out5 = out5
        return out5


    # ---- bookkeeping for the measurement contract ----------------------------------------------------------
    def launches_per_step(self, plan) -> int:
        """Kernels of this library launched by one step: the plan's forward + backward, the fused loss, AdamW and the
        workspace clear (memset)."""
        return int(plan.launches) + 3

    def launch_mode(self) -> str:
        return "eager launches through the C ABI (one ctypes call per kernel)"


def broadcast_parameters(model, src=0, process_group=None):
    """Replicas start from rank `src`'s parameters and buffers (one flat broadcast)."""
    from .params import FlatParams
    flat = FlatParams.of(model.parameters())
    dist.broadcast(flat.master, src, group=process_group)
    for b in model.buffers():
        dist.broadcast(b, src, group=process_group)
    flat.refresh_shadow(force=True)
    return flat
