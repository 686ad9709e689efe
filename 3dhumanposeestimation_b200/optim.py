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
        # the remaining keys are torch.optim.AdamW's own defaults: a state_dict saved here loads into the reference's
        # torch.optim.AdamW (main.py:130-134) and the other way round
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=True))
        self.grad_scale = grad_scale
        self._flat = None
        self._ranges = []
        self._step = 0
        self.exp_avg = self.exp_avg_sq = None
        self.grad16 = None     # set by Trainer (bf16 gradient exchange): the step reads the all-reduced bf16 gradient
        self.device_step = False   # set by Trainer (graph replay): the kernel reads the step count from the bound pose_step_state

    def _resolve(self):
        params = self.param_groups[0]["params"]
        if self._flat is None or not self._flat.intact():
            self._flat = FlatParams.of(params)
            # frozen parameters (vit_freeze_backbone, transformers.py:226-236) stay in the flat buffer but are never
            # touched: torch.optim.AdamW skips parameters without a gradient (no update, no weight decay).  The step runs
            # over the maximal contiguous trainable ranges (one launch when nothing is frozen).
            self._ranges = []
            ends = list(self._flat.offsets[1:]) + [self._flat.numel]          # padded extents (offsets are aligned)
            for p, off, end in zip(self._flat.params, self._flat.offsets, ends):
                if not p.requires_grad:
                    continue
                if self._ranges and self._ranges[-1][1] == off:
                    self._ranges[-1][1] = end
                else:
                    self._ranges.append([off, end])
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
        for lo, hi in self._ranges:
            hyper = (hi - lo, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                     float(g["weight_decay"]), 0 if self.device_step else self._step, float(self.grad_scale), 1,
                     _lib.stream_ptr())
            if self.grad16 is not None:
                code = _lib.lib().pose_adamw_step_g16(flat.master.data_ptr() + 4 * lo, flat.grad.data_ptr() + 4 * lo,
                                                      self.grad16.data_ptr() + 2 * lo, self.exp_avg.data_ptr() + 4 * lo,
                                                      self.exp_avg_sq.data_ptr() + 4 * lo, flat.shadow.data_ptr() + 2 * lo,
                                                      *hyper)
            else:
                code = _lib.lib().pose_adamw_step(flat.master.data_ptr() + 4 * lo, flat.grad.data_ptr() + 4 * lo,
                                                  self.exp_avg.data_ptr() + 4 * lo, self.exp_avg_sq.data_ptr() + 4 * lo,
                                                  flat.shadow.data_ptr() + 2 * lo, *hyper)
            _lib.check(code, "pose_adamw_step")
        flat.generation += 1          # the kernel wrote the parameters through raw pointers: _version did not move
        flat.mark_shadow_current()    # ... and refreshed the bf16 shadow itself
        return loss

    # ---- checkpointing: torch.optim.AdamW's per-parameter layout (src/train.py:300-306, main.py:130-134) ----------
    def state_dict(self):
        """``{"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}`` exactly as torch.optim.AdamW
        writes it: the flat moment buffers are scattered into per-parameter tensors (copies)."""
        flat = self._resolve()
        params = self.param_groups[0]["params"]
        self.state.clear()
        if self._step > 0:
            for p, off in zip(flat.params, flat.offsets):
                if not p.requires_grad:          # torch.optim.AdamW keeps no state for parameters without gradients
                    continue
                n = p.numel()
                self.state[p] = {"step": torch.tensor(float(self._step), dtype=torch.float32),
                                 "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                                 "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        assert len(params) == len(flat.params)
        sd = super().state_dict()
        self.state.clear()
        return sd

    def load_state_dict(self, state_dict):
        """Accepts a state dict of torch.optim.AdamW (a reference checkpoint's ``optimizer_state_dict``) or of this
        class: per-parameter moments are gathered into the flat buffers, the step count resumes bias correction."""
        flat = self._resolve()
        groups = state_dict.get("param_groups", [])
        if len(groups) != 1:
            raise ValueError("the fused AdamW holds one parameter group (as the reference does)")
        g = groups[0]
        if g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("amsgrad / maximize are not implemented by the fused AdamW")
        super().load_state_dict(state_dict)
        steps = set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for p, off in zip(flat.params, flat.offsets):
            st = self.state.get(p)
            if not st:
                continue
            n = p.numel()
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): one fused launch applies one bias "
                             "correction")
        self._step = steps.pop() if steps else 0
        self.state.clear()

    def zero_grad(self, set_to_none=True):
        """step() already cleared the flat gradient buffer; without a preceding step, clear it here.  The views
        stay attached (set_to_none would drop them)."""
        flat = self._resolve()
        flat.grad.zero_()
        if not flat.grads_attached():
            flat.attach_grads()
