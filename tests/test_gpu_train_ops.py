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


def _bn_fuse(pose, N, count, gamma, beta, rm, rv, eps=1e-5, momentum=0.1):
    f = pose._lib.PoseBnFuse()
    part = torch.empty(4 * 148 * 2 * N, device=DEV)
    mr, ss = torch.empty(2 * N, device=DEV), torch.empty(2 * N, device=DEV)
    f.partials, f.cap_floats = part.data_ptr(), part.numel()
    f.gamma, f.beta, f.eps, f.momentum, f.count = gamma.data_ptr(), beta.data_ptr(), eps, momentum, count
    f.mean_rstd, f.scale_shift = mr.data_ptr(), ss.data_ptr()
    f.running_mean, f.running_var = rm.data_ptr(), rv.data_ptr()
    return f, (part, mr, ss)


def _check_bn_fuse(y32, mr, ss, rm, rv, gamma, beta, eps=1e-5, momentum=0.1):
    """mean / rstd / scale / shift / running statistics of nn.BatchNorm2d in training mode over the fp32 rows y32 [M, N]
    (src/utils.py:186-187); fp32 sums in a different order: 1e-5 relative on the moments."""
    M, N = y32.shape
    mean = y32.double().mean(0)
    var = y32.double().var(0, unbiased=False)
    rstd = (var + eps).rsqrt()
    assert torch.allclose(mr[:N].double(), mean, rtol=1e-4, atol=1e-4 * y32.abs().max().item())
    assert torch.allclose(mr[N:].double(), rstd, rtol=2e-4)
    a = gamma.double() * rstd
    assert torch.allclose(ss[:N].double(), a, rtol=2e-4, atol=1e-6)
    assert torch.allclose(ss[N:].double(), beta.double() - mean * a, rtol=2e-4, atol=2e-4 * (mean * a).abs().max().item() + 1e-6)
    assert torch.allclose(rm.double(), momentum * mean, rtol=1e-4, atol=1e-5 * y32.abs().max().item())
    assert torch.allclose(rv.double(), (1 - momentum) * 1.0 + momentum * var * M / (M - 1), rtol=2e-4)


@pytest.mark.parametrize("M,N,K", [(128 * 64 * 4, 64, 64), (5000, 128, 128), (32768, 3072, 512), (8192 + 77, 768, 256),
                                   (300, 32, 64), (16384, 512, 3072), (40000, 256, 768)])
def test_gemm_epilogue_emits_batchnorm_statistics_from_the_accumulators(pose, M, N, K):
    """pose_bn_fuse: the 1x1-convolution GEMM of ConvBnAct (cnn.py:122-139) returns the batch statistics of its fp32
    accumulators and the folded scale / shift in the same call (128 x {32,64,128,256} tiles, ragged M, several column tiles
    per CTA grid), deterministically."""
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).bfloat16()
    w[: N // 2] += 0.05                                      # non-zero channel means
    gamma, beta = (torch.rand(N, generator=g) + 0.5).to(DEV), torch.randn(N, generator=g).to(DEV)
    runs = []
    for _ in range(2):
        rm, rv = torch.zeros(N, device=DEV), torch.ones(N, device=DEV)
        y = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        f, (part, mr, ss) = _bn_fuse(pose, N, M, gamma, beta, rm, rv)
        e = _epi(pose, y)
        e.bn = C.pointer(f)
        pose._lib.check(pose._lib.lib().pose_gemm_bf16_ex(x.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e),
                                                          pose._lib.stream_ptr()), "gemm")
        torch.cuda.synchronize()
        runs.append((y.clone(), mr.clone(), ss.clone(), rm.clone(), rv.clone()))
    y32 = x.float() @ w.float().t()
    y, mr, ss, rm, rv = runs[0]
    assert torch.allclose(y.float(), y32, rtol=2 ** -7, atol=1e-2)
    _check_bn_fuse(y32, mr, ss, rm, rv, gamma, beta)
    for a, b in zip(runs[0], runs[1]):
        assert torch.equal(a, b)                             # fixed-order partial sums: bit-reproducible


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,dil", [(4, 64, 64, 64, 3, 1, 1), (3, 64, 64, 64, 5, 2, 1), (8, 16, 512, 512, 3, 1, 6),
                                                      (5, 16, 64, 128, 3, 1, 18)])
def test_conv_epilogue_emits_batchnorm_statistics(pose, B, H, Cin, Cout, k, stride, dil):
    """The same fused statistics behind the implicit-GEMM convolution (conv1.0 / conv1.1 / WASP branches)."""
    g = torch.Generator().manual_seed(B * H + Cout + k)
    x = torch.randn(B, H, H, Cin, generator=g).to(DEV).bfloat16()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(DEV).bfloat16()
    pad = dil * (k - 1) // 2
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    M = B * Ho * Ho
    gamma, beta = (torch.rand(Cout, generator=g) + 0.5).to(DEV), torch.randn(Cout, generator=g).to(DEV)
    rm, rv = torch.zeros(Cout, device=DEV), torch.ones(Cout, device=DEV)
    y = torch.empty(M, Cout, device=DEV, dtype=torch.bfloat16)
    f, (part, mr, ss) = _bn_fuse(pose, Cout, M, gamma, beta, rm, rv)
    e = _epi(pose, y)
    e.bn = C.pointer(f)
    pose._lib.check(pose._lib.lib().pose_conv2d_bf16(x.data_ptr(), B, H, H, Cin, w.data_ptr(), Cout, k, k, stride, dil, pad,
                                                     C.byref(e), pose._lib.stream_ptr()), "conv")
    y32 = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=stride, padding=pad, dilation=dil)
    y32 = y32.permute(0, 2, 3, 1).reshape(M, Cout)
    assert torch.allclose(y.float(), y32, rtol=2 ** -7, atol=1e-2)
    _check_bn_fuse(y32, mr, ss, rm, rv, gamma, beta)
