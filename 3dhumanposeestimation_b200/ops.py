"""Thin Python wrappers over the dense-contraction entry points of the C ABI (tcgen05 GEMM)."""
from __future__ import annotations

import torch

from . import _lib
from .utils import activation_id



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
    # the cache entry lives ON the parameter object (an id()-keyed table can hand a freed parameter's entry to a new one
    # that reuses the id and the allocation)
    ent = getattr(param, "_pose_bf16", None)
    flat = getattr(param, "_pose_flat", None)
    version = param._version + (flat[0].generation if flat is not None else 0)
    if ent is None or ent[0] != version or ent[1] != param.data_ptr():
        ent = (version, param.data_ptr(), to_bf16(param.detach().contiguous()))
        param._pose_bf16 = ent
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


def _epilogue(out, ldc, bias=None, act=0, residual=None, ldr=0, preact=None, accumulate=0, drop=None):
    e = _lib.PoseGemmEpilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.residual = residual.data_ptr() if residual is not None else None
    e.C = out.data_ptr()
    e.ldc, e.ldr, e.act = ldc, ldr, act
    e.out_dtype = 0 if out.dtype == torch.float32 else 1
    e.out_scale, e.res_scale = 1.0, 1.0
    e.preact = preact.data_ptr() if preact is not None else None
    e.accumulate = accumulate
    if drop is not None:
        e.drop_p, e.drop_seed = drop
    return e


_head_calls = [0]


def _next_head_seed() -> int:
    """One dropout stream per training-mode call, derived from torch's seed (torch.manual_seed makes a run repeatable)."""
    _head_calls[0] += 1
    return ((torch.initial_seed() & 0xFFFFFFFF) << 20) + _head_calls[0] * 16


class _MlpHeadTrainFn(torch.autograd.Function):
    """Training-mode PoseRegressionHead.decoder (src/models/common.py:69-89): [Linear -> act -> Dropout(p)] x k -> Linear
    with autograd.  Forward: one tcgen05 GEMM per Linear whose epilogue applies bias, activation and the counter-based
    dropout mask and saves act'(u); backward: per layer one weight-gradient GEMM (dY^T . X), one column sum (bias) and one
    data-gradient GEMM whose epilogue multiplies by the saved derivative and the regenerated mask."""

    @staticmethod
    def forward(ctx, x, activation, p, seed, *wb):
        import ctypes as C
        lib = _lib.lib()
        sp = _lib.stream_ptr()
        n = len(wb) // 2
        weights, biases = wb[:n], wb[n:]
        act = activation_id(activation)
        M = x.shape[0]
        h = to_bf16(x.detach().contiguous()) if x.dtype != torch.bfloat16 else x.detach().contiguous()
        acts, ders = [h], []
        for i in range(n):
            w, b = weights[i], biases[i]
            last = i == n - 1
            N, K = w.shape
            out = torch.empty((M, N), dtype=torch.float32 if last else torch.bfloat16, device=x.device)
            der = None if last else torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
            e = _epilogue(out, N, b.detach() if b is not None else None, 0 if last else act, preact=der,
                          drop=None if (last or p <= 0.0) else (p, seed + i))
            _lib.check(lib.pose_gemm_bf16_ex(h.data_ptr(), K, cached_bf16(w).data_ptr(), K, M, N, K, C.byref(e), sp),
                       "pose_gemm_bf16_ex")
            if not last:
                acts.append(out)
                ders.append(der)
            h = out
        ctx.saved = (acts, ders, weights, biases, p, seed, x.dtype)
        return h

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C
        acts, ders, weights, biases, p, seed, x_dtype = ctx.saved
        lib = _lib.lib()
        sp = _lib.stream_ptr()
        n = len(weights)
        dout = dout.contiguous().float()
        M, n_out = dout.shape
        ld = (n_out + 7) // 8 * 8
        dy = torch.empty((M, ld), dtype=torch.bfloat16, device=dout.device)
        _lib.check(lib.pose_cast_f32_bf16_2d(dout.data_ptr(), n_out, M, n_out, dy.data_ptr(), ld, sp), "pose_cast_f32_bf16_2d")
        gw, gb = [None] * n, [None] * n
        dx = None
        for i in range(n - 1, -1, -1):
            w, b = weights[i], biases[i]
            N, K = w.shape
            xin = acts[i]
            if ctx.needs_input_grad[4 + i]:
                g = torch.zeros((N, K), dtype=torch.float32, device=dout.device)
                e = _epilogue(g, K, accumulate=1)
                _lib.check(lib.pose_gemm_bf16_tr(dy.data_ptr(), ld, 1, xin.data_ptr(), K, 1, N, K, M, 1, C.byref(e), sp),
                           "pose_gemm_bf16_tr")
                gw[i] = g
            if b is not None and ctx.needs_input_grad[4 + n + i]:
                g = torch.zeros((N,), dtype=torch.float32, device=dout.device)
                _lib.check(lib.pose_colsum_bf16(dy.data_ptr(), M, N, ld, g.data_ptr(), sp), "pose_colsum_bf16")
                gb[i] = g
            if i == 0 and not ctx.needs_input_grad[0]:
                break
            dx = torch.empty((M, K), dtype=torch.bfloat16, device=dout.device)
            der = ders[i - 1] if i > 0 else None
            e = _epilogue(dx, K, act=5 if der is not None else 0, residual=der, ldr=K,
                          drop=(p, seed + i - 1) if (i > 0 and p > 0.0) else None)
            _lib.check(lib.pose_gemm_bf16_tr(dy.data_ptr(), ld, 0, cached_bf16(w).data_ptr(), K, 1, M, K, N, 1, C.byref(e), sp),
                       "pose_gemm_bf16_tr")
            dy, ld = dx, K
        gx = dx.to(x_dtype) if (ctx.needs_input_grad[0] and dx is not None) else None
        return (gx, None, None, None, *gw, *gb)


def mlp_head_forward(x: torch.Tensor, linears, activation: str, dropout_p: float = 0.0) -> torch.Tensor:
    """PoseRegressionHead.decoder (src/models/common.py:69-81): [Linear -> act -> Dropout] x k -> Linear.
    Hidden activations stay bf16; the final layer writes fp32.  With gradients enabled (or dropout active) the call goes
    through _MlpHeadTrainFn so that ``loss.backward()`` reaches the head's parameters and its input."""
    _lib.require_cuda(x, "x")
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(l.weight.requires_grad for l in linears))
    if dropout_p > 0.0 or needs_grad:
        if dropout_p >= 1.0:
            raise ValueError("dropout probability must be < 1")
        if any(l.bias is None for l in linears):
            raise NotImplementedError("PoseRegressionHead layers carry a bias (nn.Linear default)")
        seed = _next_head_seed() if dropout_p > 0.0 else 0
        return _MlpHeadTrainFn.apply(x, activation, float(dropout_p), seed, *[l.weight for l in linears],
                                     *[l.bias for l in linears])
    h = to_bf16(x.contiguous()) if x.dtype != torch.bfloat16 else x.contiguous()
    n = len(linears)
    for i, lin in enumerate(linears):
        last = i == n - 1
        h = gemm_bf16(h, cached_bf16(lin.weight), lin.bias.detach() if lin.bias is not None else None,
                      act=None if last else activation, out_dtype=torch.float32 if last else torch.bfloat16)
    return h
