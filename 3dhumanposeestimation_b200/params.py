"""Flat parameter storage for training on B200.

The reference trains with ``torch.optim.AdamW(model.parameters(), lr, weight_decay)`` and
``loss.backward()`` (main.py:154-156, src/train.py:89-119).  Here every ``nn.Parameter`` of a model keeps
its reference name, shape and fp32 dtype (checkpoints stay interchangeable) but becomes a VIEW into one
flat fp32 master buffer; ``.grad`` is a view into a flat fp32 gradient buffer and the bf16 copies the
tensor-core kernels read are views into a flat bf16 shadow buffer.  That turns the optimizer step into
one kernel (pose_adamw_step), the gradient all-reduce into a few large contiguous buckets, and the
shadow refresh into one cast -- with no per-parameter launches.
"""
from __future__ import annotations

import torch

from . import _lib

_ALIGN = 64  # elements: 256 B in fp32, 128 B in bf16 (TMA needs 16 B)


class FlatParams:
    def __init__(self, params):
        params = list(params)
        if not params:
            raise ValueError("no parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatParams needs CUDA parameters (there is no CPU path): call model.cuda() first")
        self.params = params
        self.offsets = []
        n = 0
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("parameters must be fp32 on one CUDA device")
            self.offsets.append(n)
            n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = n
        self.master = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        self.index = {}
        self._views = {}
        with torch.no_grad():
            for p, off in zip(params, self.offsets):
                view = self.master[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p._pose_flat = (self, off)
                self.index[id(p)] = off
        self.attach_grads()
        self._shadow_version = None
        self.generation = 0     # bumped by every writer that bypasses torch (pose_adamw_step): caches key on it
        self.refresh_shadow()

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def of(params) -> "FlatParams":
        """The FlatParams that owns `params` (flattening them on first use)."""
        params = list(params)
        ent = getattr(params[0], "_pose_flat", None)
        if ent is not None:
            flat = ent[0]
            if len(flat.params) == len(params) and all(a is b for a, b in zip(flat.params, params)) and flat.intact():
                return flat
        return FlatParams(params)

    def intact(self) -> bool:
        """False once something re-pointed p.data (model.to(), .half(), ...): the views must be rebuilt."""
        base = self.master.data_ptr()
        return all(p.data_ptr() == base + off * 4 for p, off in zip(self.params, self.offsets))

    def attach_grads(self):
        for p, off in zip(self.params, self.offsets):
            if p.requires_grad:
                p.grad = self.grad[off:off + p.numel()].view(p.shape)

    def grads_attached(self) -> bool:
        """torch's optimizer.zero_grad(set_to_none=True) drops the views; the next backward zeroes the flat buffer
        and re-attaches them (= a new accumulation window)."""
        base = self.grad.data_ptr()
        for p, off in zip(self.params, self.offsets):
            if p.requires_grad and (p.grad is None or p.grad.data_ptr() != base + off * 4):
                return False
        return True

    # views ------------------------------------------------------------------------------------------
    # The accessors below are called several times per launch by the plans: the views are cached (the flat buffers are
    # allocated once, so a view never goes stale) -- building two or three torch views per call was a measurable share of
    # the host time of a 500-launch step.
    def _cached(self, kind, p, rows, make):
        key = (kind, id(p), None if rows is None else (int(rows[0]), int(rows[1])))
        v = self._views.get(key)
        if v is None:
            v = self._views[key] = make()
        return v

    def w16(self, p, rows=None):
        """bf16 shadow of parameter p as a 2-D [out, in...] matrix (optionally a row range)."""
        def make():
            off = self.index[id(p)]
            v = self.shadow[off:off + p.numel()].view(p.shape[0], -1)
            return v if rows is None else v[rows[0]:rows[1]]
        return self._cached(0, p, rows, make)

    def f32(self, p):
        def make():
            off = self.index[id(p)]
            return self.master[off:off + p.numel()].view(p.shape)
        return self._cached(1, p, None, make)

    def g32(self, p, rows=None):
        def make():
            off = self.index[id(p)]
            v = self.grad[off:off + p.numel()].view(p.shape)
            if rows is None:
                return v
            return v.view(p.shape[0], -1)[rows[0]:rows[1]] if p.dim() > 1 else v[rows[0]:rows[1]]
        return self._cached(2, p, rows, make)

    # shadow -----------------------------------------------------------------------------------------
    def version(self):
        return sum(p._version for p in self.params) + self.generation

    def refresh_shadow(self, force=False):
        """bf16 shadow <- fp32 master (one cast launch) whenever any parameter changed in place
        (load_state_dict, a torch optimizer); pose_adamw_step refreshes it itself."""
        v = self.version()
        if force or v != self._shadow_version:
            _lib.check(_lib.lib().pose_cast_f32_bf16(self.master.data_ptr(), self.shadow.data_ptr(), self.numel,
                                                     _lib.stream_ptr()), "pose_cast_f32_bf16")
            self._shadow_version = v

    def mark_shadow_current(self):
        self._shadow_version = self.version()
