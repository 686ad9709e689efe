"""AdamW as one fused kernel over the flat parameter buffer.

Reference: ``torch.optim.AdamW(model.parameters(), lr=LEARNING_RATE, weight_decay=WEIGHT_DECAY)`` (main.py:154-156),
stepped every ``accumulation_steps`` batches and followed by ``optimizer.zero_grad()`` (src/train.py:117-119).
Same constructor arguments and ``step()`` / ``zero_grad()`` / ``state_dict()`` surface; the update itself is
pose_adamw_step: 28 B/param of HBM traffic in one launch, fused with the refresh of the bf16 shadow weights and
with the gradient clear.
"""
from __future__ import annotations

import torch

from . import _lib
from .params import FlatParams


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        params = list(params)
        if params and isinstance(params[0], dict):
            raise NotImplementedError("parameter groups: the fused AdamW updates one flat buffer with one setting "
                                      "(the reference uses a single group)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = grad_scale
        self._flat = None
        self._step = 0
        self.exp_avg = self.exp_avg_sq = None

    def _resolve(self):
        params = self.param_groups[0]["params"]
        if self._flat is None or not self._flat.intact():
            if any(not p.requires_grad for p in params):
                raise NotImplementedError("frozen parameters (vit_freeze_backbone) are not supported by the fused AdamW")
            self._flat = FlatParams.of(params)
            if self.exp_avg is None or self.exp_avg.numel() != self._flat.numel:
                self.exp_avg = torch.zeros_like(self._flat.master)
                self.exp_avg_sq = torch.zeros_like(self._flat.master)
        return self._flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        flat = self._resolve()
        g = self.param_groups[0]
        self._step += 1
        if not flat.grads_attached():
            raise RuntimeError("gradients are not views of the flat buffer: call backward() on a model output first")
        code = _lib.lib().pose_adamw_step(flat.master.data_ptr(), flat.grad.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), flat.shadow.data_ptr(), flat.numel, float(g["lr"]),
                                          float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                          float(g["weight_decay"]), self._step, float(self.grad_scale), 1,
                                          _lib.stream_ptr())
        _lib.check(code, "pose_adamw_step")
        flat.mark_shadow_current()
        return loss

    def zero_grad(self, set_to_none=True):
        """step() already cleared the flat gradient buffer; without a preceding step, clear it here.  The views
        stay attached (set_to_none would drop them)."""
        flat = self._resolve()
        flat.grad.zero_()
        if not flat.grads_attached():
            flat.attach_grads()
