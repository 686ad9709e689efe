"""ComprehensivePoseLoss backed by one fused sm_100a forward+backward kernel.

Reference: src/loss.py:11-85.  Same constructor arguments, same ``forward(pred_joints, gt_joints) ->
(total_loss, loss_components)`` contract (dict keys ``mse_loss, l1_loss, inter_joint_loss,
abs_root_loss, total_loss``; values are 0-dim tensors).  The five scalars live in one device buffer,
so a training loop can fetch them with a single D2H copy instead of the reference's six ``.item()``
syncs (src/train.py:95-111).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .config import ABS_ROOT_LOSS_WEIGHT, INTER_JOINT_LOSS_WEIGHT, L1_LOSS_WEIGHT, MSE_LOSS_WEIGHT

_KEYS = ("mse_loss", "l1_loss", "inter_joint_loss", "abs_root_loss", "total_loss")
_workspaces = {}


def _workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)  # zero before first use
        _workspaces[key] = ws
    return ws


def pose_loss_fwd_bwd(pred, gt, weights, want_grad=True, grad_scale=1.0):
    """Raw call: returns (out5 [5] fp32 device tensor, grad [B,J,3] or None)."""
    _lib.require_cuda(pred, "pred_joints", torch.float32)
    _lib.require_cuda(gt, "gt_joints", torch.float32)
    if pred.shape != gt.shape or pred.dim() != 3 or pred.shape[2] != 3:
        raise ValueError(f"expected matching [B, J, 3] tensors, got {tuple(pred.shape)} and {tuple(gt.shape)}")
    B, J = pred.shape[0], pred.shape[1]
    lib = _lib.lib()
    nbytes = lib.pose_loss_workspace_bytes(B, J)
    ws = _workspace(pred.device, nbytes)
    out5 = torch.empty(5, dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    w = (C.c_float * 4)(*[float(x) for x in weights])
    code = lib.pose_loss_fwd_bwd(pred.data_ptr(), gt.data_ptr(), B, J, w, out5.data_ptr(),
                                 grad.data_ptr() if want_grad else None, float(grad_scale), ws.data_ptr(),
                                 ws.numel(), _lib.stream_ptr())
    _lib.check(code, "pose_loss_fwd_bwd")
    return out5, grad


class _PoseLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, weights):
        out5, grad = pose_loss_fwd_bwd(pred.detach(), gt.detach(), weights, want_grad=pred.requires_grad)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(out5)
        return out5[4].clone(), out5

    @staticmethod
    def backward(ctx, g_total, _g_out5):
        (grad,) = ctx.saved_tensors
        return grad * g_total, None, None


class ComprehensivePoseLoss(nn.Module):
    def __init__(self, l1_weight=L1_LOSS_WEIGHT, mse_weight=MSE_LOSS_WEIGHT,
                 inter_joint_loss_weight=INTER_JOINT_LOSS_WEIGHT, abs_root_loss_weight=ABS_ROOT_LOSS_WEIGHT):
        super().__init__()
        self.l1_weight = l1_weight
        self.mse_weight = mse_weight
        self.inter_joint_loss_weight = inter_joint_loss_weight
        self.abs_root_loss_weight = abs_root_loss_weight

    def _weights(self):
        return (self.mse_weight, self.l1_weight, self.inter_joint_loss_weight, self.abs_root_loss_weight)

    def forward(self, pred_joints, gt_joints):
        pred = pred_joints if pred_joints.dtype == torch.float32 else pred_joints.float()
        total, out5 = _PoseLossFn.apply(pred.contiguous(), gt_joints.contiguous(), self._weights())
        comps = {k: out5[i] for i, k in enumerate(_KEYS[:4])}
        comps["total_loss"] = total
        return total, comps
