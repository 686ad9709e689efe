"""GPU: the CNN training step -- batch-statistics BatchNorm forward / backward, convolution weight gradient as an
implicit GEMM, depthwise backward and the fused data gradient by flipped weights against plain PyTorch fp32 references
of the same ops; then one whole training step of CNNPoseEstimation (loss, every parameter's gradient, BatchNorm running
statistics) against fp32 autograd over the oracle restatement, which oracle/gen_golden.py pins to the live reference."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib(pose):
    return pose._lib.lib(), pose._lib.stream_ptr, pose._lib.check


@pytest.mark.parametrize("M,Cc,act", [(5000, 64, 2), (777, 384, 0), (300, 16, 2), (4096, 3072, 2), (64, 512, 2)])
def test_batchnorm_training_forward_backward(pose, M, Cc, act):
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(M + Cc)
    y = (torch.randn(M, Cc, generator=g) * 1.5 + 0.3).to(DEV).bfloat16()
    gamma, beta = (torch.rand(Cc, generator=g) + 0.5).to(DEV), (torch.randn(Cc, generator=g) * 0.2).to(DEV)
    rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    res = torch.randn(M, Cc, generator=g).to(DEV).bfloat16()
    da = torch.randn(M, Cc + 8, generator=g).to(DEV).bfloat16()          # column slice of a wider gradient
    part = torch.empty(1 << 20, device=DEV)                  # per-block partial sums (no atomics, fixed-order fold)
    mr, ss = torch.empty(2 * Cc, device=DEV), torch.empty(2 * Cc, device=DEV)
    out = torch.empty(M, Cc, device=DEV, dtype=torch.bfloat16)
    check(lib.pose_bn_stats_bf16(y.data_ptr(), M, Cc, Cc, part.data_ptr(), part.numel(), sp()), "stats")
    check(lib.pose_bn_finalize(part.data_ptr(), part.numel(), M, gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1, Cc, mr.data_ptr(),
                               ss.data_ptr(), rm.data_ptr(), rv.data_ptr(), sp()), "finalize")
    check(lib.pose_bn_apply_bf16(y.data_ptr(), M, Cc, ss.data_ptr(), act, 1.0, res.data_ptr(), Cc, out.data_ptr(), Cc, sp()), "apply")
    yr = y.float().requires_grad_()
    gr, br = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    rm2, rv2 = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    z = F.batch_norm(yr, rm2, rv2, gr, br, True, 0.1, 1e-5)
    a = (F.silu(z) if act == 2 else z) + res.float()
    assert torch.allclose(out.float(), a, rtol=2e-2, atol=3e-2), (out.float() - a).abs().max().item()
    assert torch.allclose(rm, rm2, rtol=1e-3, atol=1e-4) and torch.allclose(rv, rv2, rtol=1e-3, atol=1e-4)
    a.backward(da[:, 8:].float())
    dy = torch.empty(M, Cc, device=DEV, dtype=torch.bfloat16)
    coef = torch.empty(2 * Cc, device=DEV)
    dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    check(lib.pose_bn_bwd_bf16(da.data_ptr() + 16, Cc + 8, y.data_ptr(), M, Cc, ss.data_ptr(), mr.data_ptr(), act, 1.0,
                               part.data_ptr(), part.numel(), coef.data_ptr(), dy.data_ptr(), dg.data_ptr(), db.data_ptr(),
                               sp()), "bn_bwd")
    scale = yr.grad.abs().max().item()
    assert (dy.float() - yr.grad).abs().max().item() < 2e-2 * scale + 1e-3
    assert torch.allclose(dg, gr.grad, rtol=2e-2, atol=2e-2 * M ** 0.5)
    assert torch.allclose(db, br.grad, rtol=2e-2, atol=2e-2 * M ** 0.5)


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,dil", [
    (2, 64, 64, 64, 5, 2, 1),      # conv1.0 with the operand padded to 64 channels
    (2, 64, 32, 64, 5, 2, 1),      # conv1.0 as trained: 21 input channels padded to 32 (64-byte pixels, SWIZZLE_64B operand)
    (3, 32, 32, 64, 3, 1, 1),      # 32-channel pixels, 3 x 3
    (2, 64, 64, 64, 3, 1, 1),      # conv1.1
    (3, 16, 512, 512, 3, 1, 6),    # WASP dilated
    (2, 16, 128, 192, 3, 1, 18),   # dilation larger than the map
    (4, 32, 256, 512, 1, 2, 1),    # DualPath shortcut: 1x1 stride 2
    (5, 8, 64, 64, 3, 1, 1),       # ragged image count in the last 64-pixel patch
])
def test_conv_weight_gradient_implicit_gemm(pose, B, H, Cin, Cout, k, stride, dil):
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(B * 10 + H + k)
    pad = (k - 1) // 2 * dil
    x = torch.randn(B, H, H, Cin, generator=g).to(DEV).bfloat16()
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    dy = torch.randn(B, Ho, Ho, Cout, generator=g).to(DEV).bfloat16()
    w = torch.zeros(Cout, Cin, k, k, device=DEV, requires_grad=True)
    out = F.conv2d(x.float().permute(0, 3, 1, 2), w, None, stride, pad, dil)
    out.backward(dy.float().permute(0, 3, 1, 2))
    dwk = torch.full((Cout, k, k, Cin), 0.25, device=DEV)           # accumulates
    check(lib.pose_conv2d_wgrad_bf16(dy.data_ptr(), x.data_ptr(), B, H, H, Cin, Cout, k, k, stride, dil, pad, dwk.data_ptr(), 3,
                                     sp()), "wgrad")
    want = w.grad.permute(0, 2, 3, 1) + 0.25
    tol = 3e-3 * (B * Ho * Ho) ** 0.5
    assert torch.allclose(dwk, want, rtol=2e-3, atol=tol), (dwk - want).abs().max().item()


def test_param_repack_and_flipped_data_gradient(pose):
    """kind 1 / 2 / 5 of pose_param_repack; dX of a dilated 3x3 convolution = forward kernel over dY with flipped weights."""
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(3)
    B, H, Ci, Co, k, dil = 2, 16, 128, 192, 3, 6
    w = (torch.randn(Co, Ci, k, k, generator=g) / (Ci * 9) ** 0.5).to(DEV)
    tab = (pose._lib.PoseRepackEntry * 3)()
    n = Co * Ci * k * k
    for i, (dst, kind) in enumerate(((0, 1), (n, 2), (0, 5))):
        tab[i].src, tab[i].dst, tab[i].kind, tab[i].d0, tab[i].d1, tab[i].d2, tab[i].d3 = 0, dst, kind, Co, Ci, k, Ci
    tdev = torch.from_numpy(np.frombuffer(bytes(tab), dtype=np.uint8).copy()).to(DEV)
    pk16 = torch.zeros(2 * n, device=DEV, dtype=torch.bfloat16)
    check(lib.pose_param_repack(tdev.data_ptr(), 2, w.data_ptr(), None, pk16.data_ptr(), sp()), "repack")
    assert torch.equal(pk16[:n].view(Co, k, k, Ci), w.permute(0, 2, 3, 1).bfloat16())
    assert torch.equal(pk16[n:].view(Ci, k, k, Co), w.flip(2, 3).permute(1, 2, 3, 0).bfloat16())
    dy = torch.randn(B, H, H, Co, generator=g).to(DEV).bfloat16()
    dx = torch.empty(B, H, H, Ci, device=DEV, dtype=torch.float32)
    e = pose._lib.PoseGemmEpilogue()
    e.C, e.ldc, e.out_dtype, e.out_scale = dx.data_ptr(), Ci, 0, 1.0
    check(lib.pose_conv2d_bf16(dy.data_ptr(), B, H, H, Co, pk16[n:].data_ptr(), Ci, k, k, 1, dil, dil, C.byref(e), sp()), "dgrad")
    xr = torch.zeros(B, Ci, H, H, device=DEV, requires_grad=True)
    F.conv2d(xr, w.bfloat16().float(), None, 1, dil, dil).backward(dy.float().permute(0, 3, 1, 2))
    assert torch.allclose(dx, xr.grad.permute(0, 2, 3, 1), rtol=2e-3, atol=2e-2)
    # kind 5: KRSC staging -> += parameter layout
    stage = torch.randn(Co, k, k, Ci, generator=g).to(DEV)
    grad = torch.ones(Co, Ci, k, k, device=DEV)
    check(lib.pose_param_repack(tdev.data_ptr() + 2 * C.sizeof(pose._lib.PoseRepackEntry), 1, stage.data_ptr(), grad.data_ptr(),
                                None, sp()), "unpack")
    assert torch.allclose(grad, 1.0 + stage.permute(0, 3, 1, 2))


@pytest.mark.parametrize("B,H,Cc,stride", [(2, 32, 64, 1), (3, 16, 384, 2), (2, 17, 128, 2)])
def test_depthwise_backward(pose, B, H, Cc, stride):
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(H + Cc)
    x = torch.randn(B, H, H, Cc, generator=g).to(DEV).bfloat16()
    w = (torch.randn(Cc, 1, 3, 3, generator=g) * 0.3).to(DEV)
    Ho = (H - 1) // stride + 1
    dy = torch.randn(B, Ho, Ho, Cc, generator=g).to(DEV).bfloat16()
    add = torch.randn(B, H, H, Cc, generator=g).to(DEV).bfloat16()
    xr, wr = x.float().permute(0, 3, 1, 2).requires_grad_(), w.clone().requires_grad_()
    F.conv2d(xr, wr, None, stride, 1, groups=Cc).backward(dy.float().permute(0, 3, 1, 2))
    wd = w.view(Cc, 9).t().contiguous()
    dx = torch.empty_like(x)
    dw = torch.zeros(Cc, 1, 3, 3, device=DEV)
    check(lib.pose_dwconv3x3_bwd_bf16(dy.data_ptr(), x.data_ptr(), wd.data_ptr(), B, H, H, Cc, stride, add.data_ptr(),
                                      dx.data_ptr(), dw.data_ptr(), sp()), "dw_bwd")
    want = xr.grad.permute(0, 2, 3, 1) + add.float()
    assert torch.allclose(dx.float(), want, rtol=2e-2, atol=3e-2)
    assert torch.allclose(dw, wr.grad, rtol=2e-3, atol=2e-3 * (B * Ho * Ho) ** 0.5)


# ---------------------------------------------------------------------------------------------------------------------
def _setup(pose, golden, **kw):
    from oracle import torch_models as tm
    gd = golden("cnn_train_256.npz")
    cfg = pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256, regression_dropout=0.0, **kw)
    m = pose.CNNPoseEstimation(cfg)
    sd = tm.fill_state_dict(m.state_dict(), seed=int(gd["fill_seed"]))
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(int(gd["input_seed"]))
    img = torch.rand(4, 3, 256, 256, generator=g).to(DEV)
    dep = torch.rand(4, 1, 256, 256, generator=g).to(DEV)
    return gd, m.to(DEV), {k: v.to(DEV) for k, v in sd.items()}, tm, img, dep, torch.from_numpy(gd["kp"]).to(DEV), \
        torch.from_numpy(gd["gt"]).to(DEV)


def test_training_step_matches_fp32_autograd_and_the_reference(pose, golden):
    """A randomly initialised train-mode BatchNorm network amplifies rounding noise layer by layer (the forward
    perturbation grows with depth exactly like the well-known gradient growth of BN networks at initialisation), so a bf16
    step cannot sit within 0.5 mm of fp32 here -- PyTorch's own bf16 autocast of the same model is ~10 mm / ~20 % away.
    The bars: every kernel is tight on its own (tests above); the whole step must be at least as close to the fp32
    reference as torch.autocast(bfloat16) is, and gradient norms must agree with the live reference's."""
    gd, m, sd, tm, img, dep, kp, gt = _setup(pose, golden)
    m.train()
    crit = pose.ComprehensivePoseLoss()
    pred = m(img, dep, kp)
    total, _ = crit(pred, gt)
    total.backward()
    names = [n for n, _ in m.named_parameters()]
    iu = torch.triu_indices(17, 17, 1, device=DEV)
    pd = lambda t: torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]   # noqa: E731

    def oracle_step(autocast):
        sdg = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            po, stats = tm.cnn_forward(sdg, m.config, img, dep, kp, train=True, return_stats=True)
        po = po.float()
        d = po - gt
        (((d ** 2).mean() + d.abs().mean() + 100.0 * (pd(po) - pd(gt)).abs().mean() + d[:, 0].abs().mean())).backward()
        return po.detach(), {n: sdg[n].grad.double() for n in names}, stats

    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False     # the fp32 reference is IEEE fp32
    try:
        p32, g32, new_stats = oracle_step(False)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    p16, g16, _ = oracle_step(True)
    ref_pred = torch.from_numpy(gd["pred"]).to(DEV)
    assert pose.utils.compute_mpjpe(p32, ref_pred).item() < 0.2              # fp32 oracle on this GPU == live reference
    mpjpe = pose.utils.compute_mpjpe(pred.detach(), p32).item()
    mpjpe16 = pose.utils.compute_mpjpe(p16, p32).item()
    # as close to fp32 as torch.autocast(bfloat16).  Both numbers are samples of the same amplified rounding noise: changing
    # only the summation ORDER of the batch statistics (transposing vs TMA-store convolution epilogue, identical arithmetic)
    # moved ours between 9.0 and 9.3 mm at this batch of 4 while autocast sat at 9.1-10.3 mm over torch versions / boxes,
    # hence the 10 % band on a comparison of two noisy quantities (the B = 32 / 128 runs below keep the strict bar)
    assert mpjpe < max(0.5, 1.1 * mpjpe16), (mpjpe, mpjpe16)
    assert abs(total.item() - float(gd["loss"])) < 1e-2 * float(gd["loss"])
    gn_ref = dict(zip(list(gd["grad_names"]), gd["grad_norms"]))
    floor = 2e-6 * float(max(gd["grad_norms"]))      # analytically-zero gradients hold rounding noise (gen_golden.py)
    ours, auto, dev, dev16 = [], [], [], []
    for n, p in m.named_parameters():
        r = g32[n]
        rn = r.norm().item()
        assert abs(rn - gn_ref[n]) <= 1e-2 * gn_ref[n] + floor, (n, rn, gn_ref[n])      # oracle == live reference
        an = p.grad.double().norm().item()
        dev.append((abs(an - rn) / (rn + 50 * floor), n))
        dev16.append(abs(g16[n].norm().item() - rn) / (rn + 50 * floor))
        ours.append(((p.grad.double() - r).norm().item() / (rn + 50 * floor), n))
        auto.append((g16[n] - r).norm().item() / (rn + 50 * floor))
    mean_ours, mean_auto = sum(r for r, _ in ours) / len(ours), sum(auto) / len(auto)
    # every gradient has the right size: 1.6 % mean deviation of the norms (measured), worst parameter (an SE excitation
    # weight behind 4 samples x 64 x 64 pixels of squeeze) no further out than autocast's worst
    assert sum(d for d, _ in dev) / len(dev) < 0.03, sorted(dev, reverse=True)[:8]
    assert max(dev)[0] < max(0.15, 1.1 * max(dev16)), (sorted(dev, reverse=True)[:8], max(dev16))
    # and the right direction: mean relative error no larger than torch.autocast(bfloat16)'s
    assert mean_ours < 1.0 * mean_auto + 0.005, (mean_ours, mean_auto, sorted(ours, reverse=True)[:8])
    assert max(ours)[0] < 1.25 * max(auto) + 0.02, (sorted(ours, reverse=True)[:8], max(auto))
    # BatchNorm running statistics follow nn.BatchNorm2d (momentum 0.1, unbiased variance)
    after = m.state_dict()
    for k, v in new_stats.items():
        assert torch.allclose(after[k], v, rtol=2e-2, atol=2e-2), k
    assert int(after["conv1.0.norm.num_batches_tracked"]) == 1


@pytest.mark.parametrize("B", [32, 128])
def test_training_step_at_the_benchmark_batch_sizes(pose, B):
    """BASELINE configs[2] runs batch 128 per GPU: the whole step (train-mode forward with batch statistics, composite loss,
    backward) against the fp32 oracle's autograd on the same B200, next to torch.autocast(bfloat16) of the same model.
    Bars: MPJPE and mean relative gradient error no larger than autocast's, gradient norms within 3 % on average."""
    from oracle import torch_models as tm
    torch.manual_seed(0)
    cfg = pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256, regression_dropout=0.0)
    m = pose.CNNPoseEstimation(cfg)
    sd = tm.fill_state_dict(m.state_dict(), seed=5)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    sd = {k: v.to(DEV) for k, v in sd.items()}
    g = torch.Generator().manual_seed(100 + B)
    img, dep = torch.rand(B, 3, 256, 256, generator=g).to(DEV), torch.rand(B, 1, 256, 256, generator=g).to(DEV)
    kp = (torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05).to(DEV)
    gt = (torch.randn(B, 17, 3, generator=g) * 300).to(DEV)
    pred = m(img, dep, kp)
    total, _ = pose.ComprehensivePoseLoss()(pred, gt)
    total.backward()
    names = [n for n, _ in m.named_parameters()]

    def oracle_step(autocast):
        sdg = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            po, _ = tm.cnn_forward(sdg, cfg, img, dep, kp, train=True, return_stats=True)
        po = po.float()
        loss = tm.composite_loss(po, gt)
        loss.backward()
        out = po.detach(), {n: sdg[n].grad.double() for n in names}, loss.item()
        del sdg
        return out

    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        p32, g32, l32 = oracle_step(False)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    p16, g16, _ = oracle_step(True)
    torch.cuda.empty_cache()
    mpjpe = pose.utils.compute_mpjpe(pred.detach(), p32).item()
    mpjpe16 = pose.utils.compute_mpjpe(p16, p32).item()
    assert mpjpe < mpjpe16, (mpjpe, mpjpe16)
    assert abs(total.item() - l32) < 1e-2 * l32, (total.item(), l32)
    gmax = max(v.norm().item() for v in g32.values())
    ours, auto, dev = [], [], []
    for n, p in m.named_parameters():
        r = g32[n]
        rn = r.norm().item()
        if rn < 1e-5 * gmax:
            continue                 # analytically-zero gradients (a conv bias in front of a BatchNorm) hold rounding noise
        ours.append(((p.grad.double() - r).norm().item() / rn, n))
        auto.append((g16[n] - r).norm().item() / rn)
        dev.append((abs(p.grad.double().norm().item() - rn) / rn, n))
    mean_ours, mean_auto = sum(r for r, _ in ours) / len(ours), sum(auto) / len(auto)
    print(f"CNN train B={B}: MPJPE {mpjpe:.2f} mm (autocast {mpjpe16:.2f}), mean gradient error {mean_ours:.4f} (autocast {mean_auto:.4f}), "
          f"norm deviation mean {sum(d for d, _ in dev) / len(dev):.4f} max {max(dev)}")
    assert mean_ours < mean_auto + 0.005, (mean_ours, mean_auto)
    assert sum(d for d, _ in dev) / len(dev) < 0.03 and max(dev)[0] < 0.2, sorted(dev, reverse=True)[:6]


def test_trainer_reduces_the_loss(pose, golden):
    gd, m, sd, tm, img, dep, kp, gt = _setup(pose, golden)
    m.train()
    train = __import__("importlib").import_module("3dhumanposeestimation_b200.train")
    tr = train.Trainer(m, pose.ComprehensivePoseLoss(), lr=2e-4)
    losses = [tr.step(img, dep, kp, gt)[4].item() for _ in range(5)]
    assert losses[-1] < losses[0], losses


def _bn_setup(M, Cc, g):
    y = (torch.randn(M, Cc, generator=g) * 1.5 + 0.3).to(DEV).bfloat16()
    gamma, beta = (torch.rand(Cc, generator=g) + 0.5).to(DEV), (torch.randn(Cc, generator=g) * 0.2).to(DEV)
    mean, var = y.float().mean(0), y.float().var(0, unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    mr = torch.cat([mean, rstd]).contiguous()
    ss = torch.cat([gamma * rstd, beta - mean * gamma * rstd]).contiguous()
    return y, gamma, beta, mr, ss


def _bn_bwd_reference(y, gamma, beta, dA, act):
    """fp32 autograd of a = act(batch_norm(y)) for the incoming gradient dA."""
    yr = y.float().requires_grad_()
    gr, br = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    z = F.batch_norm(yr, None, None, gr, br, True, 0.1, 1e-5)
    a = F.silu(z) if act == 2 else F.relu(z)
    a.backward(dA)
    return yr.grad, gr.grad, br.grad


@pytest.mark.parametrize("B,HW,Cc,act", [(3, 200, 64, 2), (2, 1024, 768, 2), (5, 64, 3072, 1), (4, 37, 16, 2)])
def test_batchnorm_backward_reduction_carried_by_the_gate_backward(pose, B, HW, Cc, act):
    """pose_gate_bwd_apply_bn_bf16 + pose_bn_bwd_from_dz_bf16 == the SE / ECA gate backward followed by the BatchNorm + act
    backward of the layer in front of it (fp32 autograd of both)."""
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(B * 7 + Cc)
    M = B * HW
    y, gamma, beta, mr, ss = _bn_setup(M, Cc, g)
    dO = torch.randn(M, Cc, generator=g).to(DEV).bfloat16()
    gate = torch.rand(B, Cc, generator=g).to(DEV)
    dmean = torch.randn(B, Cc, generator=g).to(DEV).bfloat16()
    part = torch.empty(4 << 20, device=DEV)
    coef = torch.empty(2 * Cc, device=DEV)
    dz, dy = torch.empty_like(y), torch.empty_like(y)
    dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    n_parts = C.c_int(0)
    check(lib.pose_gate_bwd_apply_bn_bf16(dO.data_ptr(), gate.data_ptr(), dmean.data_ptr(), 1.0 / HW, B, HW, Cc, y.data_ptr(),
                                          ss.data_ptr(), act, dz.data_ptr(), part.data_ptr(), part.numel(), C.byref(n_parts), sp()),
          "gate_bn")
    assert n_parts.value > 0
    check(lib.pose_bn_bwd_from_dz_bf16(dz.data_ptr(), Cc, y.data_ptr(), M, Cc, ss.data_ptr(), mr.data_ptr(), part.data_ptr(),
                                       n_parts.value, coef.data_ptr(), dy.data_ptr(), dg.data_ptr(), db.data_ptr(), sp()), "from_dz")
    dA = (dO.float().view(B, HW, Cc) * gate[:, None, :] + dmean.float()[:, None, :] / HW).view(M, Cc)
    want, wg, wb = _bn_bwd_reference(y, gamma, beta, dA, act)
    scale = want.abs().max().item()
    assert (dy.float() - want).abs().max().item() < 2e-2 * scale + 1e-3
    assert torch.allclose(dg, wg, rtol=2e-2, atol=2e-2 * M ** 0.5)
    assert torch.allclose(db, wb, rtol=2e-2, atol=2e-2 * M ** 0.5)
    # deterministic: fixed-order partial sums
    dg2, db2 = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    check(lib.pose_gate_bwd_apply_bn_bf16(dO.data_ptr(), gate.data_ptr(), dmean.data_ptr(), 1.0 / HW, B, HW, Cc, y.data_ptr(),
                                          ss.data_ptr(), act, dz.data_ptr(), part.data_ptr(), part.numel(), C.byref(n_parts), sp()),
          "gate_bn")
    check(lib.pose_bn_bwd_from_dz_bf16(dz.data_ptr(), Cc, y.data_ptr(), M, Cc, ss.data_ptr(), mr.data_ptr(), part.data_ptr(),
                                       n_parts.value, coef.data_ptr(), dy.data_ptr(), dg2.data_ptr(), db2.data_ptr(), sp()), "from_dz")
    assert torch.equal(dg, dg2) and torch.equal(db, db2)


@pytest.mark.parametrize("B,H,W,Cc,act", [(2, 32, 32, 128, 2), (3, 16, 16, 768, 2), (2, 19, 13, 64, 1), (1, 8, 40, 24, 2)])
def test_batchnorm_backward_reduction_carried_by_the_depthwise_data_gradient(pose, B, H, W, Cc, act):
    """pose_dwconv3x3_bnbwd_bf16 + pose_bn_bwd_from_dz_bf16 == data gradient of a stride-1 depthwise 3x3 followed by the
    BatchNorm + act backward of the layer that produced its input (fp32 autograd of both)."""
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(H * 3 + Cc)
    M = B * H * W
    y, gamma, beta, mr, ss = _bn_setup(M, Cc, g)
    dyw = torch.randn(B, H, W, Cc, generator=g).to(DEV).bfloat16()            # gradient of the depthwise OUTPUT
    wdw = (torch.randn(Cc, 1, 3, 3, generator=g) * 0.3).to(DEV)
    wflip = wdw.view(Cc, 9).flip(1).t().contiguous()                          # [9, C], taps flipped (repack kind 6)
    parts = B * lib.pose_dwconv3x3_pool_parts(H, W, 1)
    part = torch.empty(max(parts * 2 * Cc, 1 << 16), device=DEV)
    coef = torch.empty(2 * Cc, device=DEV)
    dz, dy = torch.empty_like(y), torch.empty_like(y)
    dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    check(lib.pose_dwconv3x3_bnbwd_bf16(dyw.data_ptr(), B, H, W, Cc, wflip.data_ptr(), y.data_ptr(), ss.data_ptr(), act,
                                        dz.data_ptr(), part.data_ptr(), part.numel(), sp()), "dw_bnbwd")
    check(lib.pose_bn_bwd_from_dz_bf16(dz.data_ptr(), Cc, y.data_ptr(), M, Cc, ss.data_ptr(), mr.data_ptr(), part.data_ptr(), parts,
                                       coef.data_ptr(), dy.data_ptr(), dg.data_ptr(), db.data_ptr(), sp()), "from_dz")
    xin = torch.zeros(B, Cc, H, W, device=DEV, requires_grad=True)
    F.conv2d(xin, wdw, None, 1, 1, 1, Cc).backward(dyw.float().permute(0, 3, 1, 2))
    dA = xin.grad.permute(0, 2, 3, 1).reshape(M, Cc)
    want, wg, wb = _bn_bwd_reference(y, gamma, beta, dA, act)
    scale = want.abs().max().item()
    assert (dy.float() - want).abs().max().item() < 2e-2 * scale + 1e-3
    assert torch.allclose(dg, wg, rtol=2e-2, atol=2e-2 * M ** 0.5)
    assert torch.allclose(db, wb, rtol=2e-2, atol=2e-2 * M ** 0.5)
