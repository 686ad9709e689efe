"""Thin Python wrappers over the dense-contraction entry points of the C ABI (tcgen05 GEMM)."""
from __future__ import annotations

import torch

from . import _lib
from .utils import activation_id

_bf16_cache = {}


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 through pose_cast_f32_bf16 (operand preparation)."""
    if x.dtype == torch.bfloat16:
        return x
    _lib.require_cuda(x, "x", torch.float32)
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().pose_cast_f32_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _lib.stream_ptr()),
               "pose_cast_f32_bf16")
    return out


def cached_bf16(param: torch.Tensor) -> torch.Tensor:
    """bf16 shadow copy of an fp32 master parameter, refreshed when the parameter is updated in place: torch writes
    bump ``_version``; the fused AdamW writes through raw pointers and bumps ``FlatParams.generation`` instead.
    Master weights keep the reference's fp32 [out, in] layout."""
    key = id(param)
    ent = _bf16_cache.get(key)
    flat = getattr(param, "_pose_flat", None)
    version = param._version + (flat[0].generation if flat is not None else 0)
    if ent is None or ent[0] != version or ent[1] != param.data_ptr():
        ent = (version, param.data_ptr(), to_bf16(param.detach().contiguous()))
        _bf16_cache[key] = ent
    return ent[2]


def gemm_bf16(a: torch.Tensor, w: torch.Tensor, bias=None, act=None, out_dtype=torch.float32) -> torch.Tensor:
    """act(a @ w.T + bias): a [M,K] bf16, w [N,K] bf16, bias [N] fp32 -> [M,N] fp32 or bf16."""
    _lib.require_cuda(a, "a", torch.bfloat16)
    _lib.require_cuda(w, "w", torch.bfloat16)
    if a.dim() != 2 or w.dim() != 2 or a.shape[1] != w.shape[1]:
        raise ValueError(f"gemm shapes {tuple(a.shape)} x {tuple(w.shape)}^T")
    if bias is not None:
        _lib.require_cuda(bias, "bias", torch.float32)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    code = _lib.lib().pose_gemm_bf16(a.data_ptr(), K, w.data_ptr(), K, bias.data_ptr() if bias is not None else None,
                                     out.data_ptr(), N, M, N, K, activation_id(act),
                                     0 if out_dtype == torch.float32 else 1, _lib.stream_ptr())
    _lib.check(code, "pose_gemm_bf16")
    return out


def MLP_HEAD_LAUNCHES(n_linear: int) -> int:
    """Kernels launched by mlp_head_forward: one operand cast + one fused GEMM per Linear."""
    return 1 + n_linear


def mlp_head_forward(x: torch.Tensor, linears, activation: str, dropout_p: float = 0.0) -> torch.Tensor:
    """PoseRegressionHead.decoder (src/models/common.py:69-81): [Linear -> act -> Dropout] x k -> Linear.
    Hidden activations stay bf16; the final layer writes fp32."""
    if dropout_p > 0.0:
        raise NotImplementedError("training-mode dropout in the fused head is not implemented yet; call .eval() "
                                  "or construct the head with dropout=0")
    h = to_bf16(x.contiguous()) if x.dtype != torch.bfloat16 else x.contiguous()
    n = len(linears)
    for i, lin in enumerate(linears):
        last = i == n - 1
        h = gemm_bf16(h, cached_bf16(lin.weight), lin.bias.detach() if lin.bias is not None else None,
                      act=None if last else activation, out_dtype=torch.float32 if last else torch.bfloat16)
    return h
