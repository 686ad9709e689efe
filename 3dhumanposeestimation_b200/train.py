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
from .dist import SectionCoalescer, allreduce_range
from .loss import pose_loss_fwd_bwd
from .optim import AdamW


class Trainer:
    def __init__(self, model, criterion, optimizer=None, accumulation_steps=1, lr=1e-3, weight_decay=0.01,
                 process_group=None, bucket_bytes=64 << 20, grad_wire=None, graph=None):
        self.model, self.crit = model, criterion
        self.opt = optimizer if optimizer is not None else AdamW(model.parameters(), lr=lr, weight_decay=weight_decay)
        self.accum = int(accumulation_steps)
        self.micro = 0
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.pg = process_group
        self.bucket_bytes = bucket_bytes
        self.comm_stream = torch.cuda.Stream() if self.world > 1 else None
        self._pending = []
        # coalescing threshold of the gradient exchange (bytes of fp32 gradient; 0 = every announced section on its own).
        # Measured at 8 GPUs, ViT B = 64 / GPU: per-section exchange 26.08 ms / step, 128 MB 25.62, one exchange after the
        # backward 25.68 (1 GPU: 24.1): short collectives spread over the whole backward cost more in interference
        # with the persistent GEMM grids than their overlap hides
        self.comm_min_bytes = int(os.environ.get("POSE_DP_MIN_BYTES", str(128 << 20)))
        self._coalesce = SectionCoalescer(self.comm_min_bytes // 4)
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
        # CUDA-graph replay of the step (accumulation windows of one micro-batch; POSE_TRAIN_GRAPH=0 keeps eager launches)
        if graph is None:
            graph = os.environ.get("POSE_TRAIN_GRAPH", "1") != "0"
        self.use_graph = bool(graph) and self.accum == 1
        self._graphs = {}
        self._state = None
        self._state_step = None

    # ---- gradient exchange ---------------------------------------------------------------------------
    def _section_done(self, flat, lo, hi):
        """Gradients in flat[lo:hi) are final: all-reduce them on the communication stream (bucketed).  Sections arrive
        from the tail of the flat buffer towards its head; they are coalesced until `comm_min_bytes` of gradient are
        pending (or the head is reached), so that the exchange runs as a few long collectives instead of many short
        ones sharing the SMs with the backward kernels for the whole pass."""
        if self.world == 1 or hi <= lo:
            return
        self._flat_for_flush = flat
        for a, b in self._coalesce.add(lo, hi):
            self._exchange(flat, a, b)

    def _exchange(self, flat, lo, hi):
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
            for a, b in self._coalesce.flush():          # (a backward that did not announce the head of the buffer)
                self._exchange(self._flat_for_flush, a, b)
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    # ---- one batch -----------------------------------------------------------------------------------
    def step(self, images, depths, keypoints_2d, gt_joints):
        """Returns the device tensor [mse, l1, inter_joint, abs_root, total] of this batch (no host sync).  In graph mode the
        tensor is a static buffer that the next step overwrites."""
        if self.use_graph:
            return self._step_graph(images, depths, keypoints_2d, gt_joints)
        return self._step_eager(images, depths, keypoints_2d, gt_joints)

    def _step_eager(self, images, depths, keypoints_2d, gt_joints, state_mode=False):
        m = self.model
        B = images.shape[0]
        plan = m.plan(B, images.device)
        last = (self.micro + 1) % self.accum == 0
        if state_mode:
            # per-step state in device memory: the tick advances the dropout key and AdamW's step count, the kernels read them
            _lib.bind_step_state(self._state)
            _lib.check(_lib.lib().pose_step_tick(self._state.data_ptr(), _lib.stream_ptr()), "pose_step_tick")
        try:
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
                self.opt.device_step = state_mode
                self.opt.step()                                  # also clears the flat gradient buffer
        finally:
            if state_mode:
                self.opt.device_step = False
                _lib.bind_step_state(None)
        self.out5 = out5
        return out5

    # ---- the same step as ONE CUDA-graph launch ---------------------------------------------------------------
    def _sync_state(self):
        """Device copy of what the host knows: AdamW's step count (a loaded optimizer state changes it)."""
        if self._state is None:
            self._state = torch.zeros(4, dtype=torch.int32, device=next(self.model.parameters()).device)
            self._state_step = None
        if self._state_step != self.opt._step:
            self._state[1:2].fill_(int(self.opt._step))
            self._state_step = self.opt._step

    def _step_graph(self, images, depths, keypoints_2d, gt_joints):
        """Two eager steps (allocations, kernel attributes, NCCL warm-up), then the whole step -- tick, forward, loss,
        backward, gradient exchange, AdamW -- is captured once per batch shape and replayed: one launch per step instead of
        500-600, no gaps between the kernels.  What changes from step to step (dropout key, AdamW step count) is read from
        the device-resident pose_step_state."""
        inputs = (images, depths, keypoints_2d, gt_joints)
        key = tuple((tuple(t.shape), t.dtype, t.device.index) for t in inputs)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = {"warm": 0}
        self._sync_state()
        if g["warm"] < 2:
            g["warm"] += 1
            out5 = self._step_eager(*inputs, state_mode=True)
            self._state_step = self.opt._step
            return out5
        flat = self.opt._resolve()
        if "graph" not in g:
            static = [torch.empty_like(t) for t in inputs]
            for s_, t in zip(static, inputs):
                s_.copy_(t)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    out5 = self._step_eager(*static, state_mode=True)
            except Exception:
                # capture failed (an op that cannot be captured): stay eager from here on
                self.use_graph = False
                torch.cuda.synchronize()
                raise
            g.update(graph=graph, static=static, out5=out5)
            # the capture ran the host side of one step (host counters advanced) without executing it: replay executes it
            self._state_step = self.opt._step
            graph.replay()
            flat.generation += 1
            flat.mark_shadow_current()
            return out5
        for s_, t in zip(g["static"], inputs):
            if s_.data_ptr() != t.data_ptr():
                s_.copy_(t)
        flat.refresh_shadow()                 # parameters changed in place since the last step (load_state_dict, ...)
        self.opt._step += 1
        self._state_step = self.opt._step
        g["graph"].replay()
        flat.generation += 1                  # the replayed AdamW wrote the parameters and refreshed the bf16 shadow
        flat.mark_shadow_current()
        self.out5 = g["out5"]
        return g["out5"]

    # ---- bookkeeping for the measurement contract ----------------------------------------------------------
    def launches_per_step(self, plan) -> int:
        """Kernels of this library launched by one step: the plan's forward + backward, the fused loss, AdamW and the
        workspace clear (memset)."""
        return int(plan.launches) + 3

    def launch_mode(self) -> str:
        if self.use_graph:
            return "one CUDA-graph launch per step (captured from the C-ABI launches; dropout key and AdamW step in device memory)"
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
