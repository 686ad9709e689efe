"""GPU: the backward-pass building blocks (transposed-operand tcgen05 GEMMs, split-K accumulation, activation
derivative / pre-activation epilogues) against plain PyTorch fp32 references of the same ops on the same
bf16-rounded operands."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _epi(pose, out, bias=None, act=0, residual=None, out_scale=1.0, res_scale=0.0, preact=None, accumulate=0):
    e = pose._lib.PoseGemmEpilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.residual = residual.data_ptr() if residual is not None else None
    e.C = out.data_ptr()
    e.ldc = out.shape[-1]
    e.ldr = residual.shape[-1] if residual is not None else 0
    e.act = act
    e.out_dtype = 0 if out.dtype == torch.float32 else 1
    e.out_scale, e.res_scale = out_scale, res_scale
    e.preact = preact.data_ptr() if preact is not None else None
    e.accumulate = accumulate
    return e


def _tr(pose, A, a_mn, W, b_mn, M, N, K, splits, e):
    pose._lib.check(pose._lib.lib().pose_gemm_bf16_tr(A.data_ptr(), A.shape[-1], a_mn, W.data_ptr(), W.shape[-1], b_mn,
                                                      M, N, K, splits, C.byref(e), pose._lib.stream_ptr()), "gemm_tr")


@pytest.mark.parametrize("M,Nout,Kin", [(257 * 3, 768, 768), (300, 51, 256), (1000, 3072, 768), (64, 256, 512),
                                        (130, 64, 128), (16448, 768, 768), (8200, 3072, 768)])   # the last two: 128 x 256 tiles
def test_data_gradient_gemm_reads_the_weight_in_place(pose, M, Nout, Kin):
    g = torch.Generator().manual_seed(M + Nout)
    ldy = (Nout + 7) // 8 * 8
    dy = torch.zeros(M, ldy)
    dy[:, :Nout] = torch.randn(M, Nout, generator=g)
    dy = dy.to(DEV).bfloat16()
    w = (torch.randn(Nout, Kin, generator=g) / Nout ** 0.5).to(DEV).bfloat16()
    dx = torch.empty(M, Kin, device=DEV, dtype=torch.float32)
    _tr(pose, dy, 0, w, 1, M, Kin, Nout, 1, _epi(pose, dx))
    want = dy[:, :Nout].float() @ w.float()
    assert torch.allclose(dx, want, rtol=1e-3, atol=1e-3), (dx - want).abs().max().item()
    # fused activation derivative: dU = (dH . W) * act'(u); act 5 multiplies by the derivative the forward GEMM saved,
    # act 6 recomputes silu' from a saved pre-activation
    u = torch.randn(M, Kin, generator=g).to(DEV).bfloat16()
    uu = u.float().requires_grad_()
    F.gelu(uu).backward(torch.ones_like(uu))
    gp = uu.grad.bfloat16()
    du = torch.empty(M, Kin, device=DEV, dtype=torch.bfloat16)
    _tr(pose, dy, 0, w, 1, M, Kin, Nout, 1, _epi(pose, du, act=5, residual=gp))
    assert torch.allclose(du.float(), want * gp.float(), rtol=2e-2, atol=2e-2)
    du = torch.empty(M, Kin, device=DEV, dtype=torch.bfloat16)
    _tr(pose, dy, 0, w, 1, M, Kin, Nout, 1, _epi(pose, du, act=6, residual=u))
    uu = u.float().requires_grad_()
    F.silu(uu).backward(want)
    assert torch.allclose(du.float(), uu.grad, rtol=2e-2, atol=2e-2), (du.float() - uu.grad).abs().max().item()


@pytest.mark.parametrize("M,Nout,Kin,splits", [(257 * 8, 768, 768, 4), (1000, 51, 256, 3), (4096, 3072, 768, 1),
                                               (64, 1024, 768, 1), (333, 200, 72, 2),
                                               (4096, 3072, 768, 4), (16448, 768, 3072, 6)])   # the last two: 128 x 256 tiles
def test_weight_gradient_gemm_split_k_accumulates(pose, M, Nout, Kin, splits):
    g = torch.Generator().manual_seed(M + Nout + 1)
    ldy = (Nout + 7) // 8 * 8
    dy = torch.zeros(M, ldy)
    dy[:, :Nout] = torch.randn(M, Nout, generator=g)
    dy = dy.to(DEV).bfloat16()
    x = torch.randn(M, Kin, generator=g).to(DEV).bfloat16()
    dw = torch.full((Nout, Kin), 0.5, device=DEV, dtype=torch.float32)     # accumulates into an existing .grad
    _tr(pose, dy, 1, x, 1, Nout, Kin, M, splits, _epi(pose, dw, accumulate=1))
    want = 0.5 + dy[:, :Nout].float().t() @ x.float()
    tol = 2e-3 * (M ** 0.5)
    assert torch.allclose(dw, want, rtol=1e-3, atol=tol), (dw - want).abs().max().item()


@pytest.mark.parametrize("M", [500, 16448])        # 16448 rows: 128 x 256 tiles with the two-chunk TMA-store epilogue
def test_forward_gemm_saves_the_activation_derivative(pose, M):
    g = torch.Generator().manual_seed(3)
    K, N = 768, 3072
    a = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).bfloat16()
    b = torch.randn(N, generator=g).to(DEV)
    h = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    u = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    e = _epi(pose, h, bias=b, act=3, preact=u)
    pose._lib.check(pose._lib.lib().pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e),
                                                      pose._lib.stream_ptr()), "gemm")
    want_u = (a.float() @ w.float().t() + b).requires_grad_()
    F.gelu(want_u).backward(torch.ones_like(want_u))
    assert torch.allclose(u.float(), want_u.grad, rtol=2 ** -7, atol=2e-2)          # gelu'(a w^T + b), exact-erf form
    assert torch.allclose(h.float(), F.gelu(want_u.detach()), rtol=2 ** -7, atol=2e-2)
    for act, fn in ((2, F.silu), (1, F.relu)):
        e = _epi(pose, h, bias=b, act=act, preact=u)
        pose._lib.check(pose._lib.lib().pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e),
                                                          pose._lib.stream_ptr()), "gemm")
        wu = want_u.detach().clone().requires_grad_()
        fn(wu).backward(torch.ones_like(wu))
        far = (wu.detach().abs() > 0.05)          # relu' is discontinuous at 0: compare away from it
        assert torch.allclose(u.float()[far], wu.grad[far], rtol=2e-2, atol=2e-2)
