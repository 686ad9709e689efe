"""GPU: implicit-GEMM convolution, the bandwidth-bound CNN pieces and the whole eval-mode CNN forward against
plain PyTorch fp32 references of the same ops (floating-point kernels) and against the golden frozen from the
live reference.  Tolerance for the bf16 forward: 0.5 mm MPJPE (BASELINE.json:north_star)."""
import ctypes as C
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _epi(pose, out, bias=None, act=0, residual=None, out_scale=1.0, res_scale=0.0):
    e = pose._lib.PoseGemmEpilogue()
    e.bias = bias.data_ptr() if bias is not None else None
    e.residual = residual.data_ptr() if residual is not None else None
    e.C = out.data_ptr()
    e.ldc = out.shape[-1]
    e.ldr = residual.shape[-1] if residual is not None else 0
    e.act = act
    e.out_dtype = 0 if out.dtype == torch.float32 else 1
    e.out_scale, e.res_scale = out_scale, res_scale
    return e


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,dil", [
    (2, 64, 32, 64, 5, 2, 1),      # conv1.0 shape class: 5x5 stride 2, 32 padded channels (64 B swizzle)
    (2, 128, 64, 64, 3, 1, 1),     # conv1.1: 3x3 @128^2
    (3, 16, 512, 512, 3, 1, 6),    # WASP dilated
    (2, 16, 128, 96, 3, 1, 18),    # dilation larger than the map: mostly padding
    (4, 32, 256, 512, 1, 2, 1),    # DualPath shortcut: 1x1 stride 2
    (3, 4, 64, 64, 3, 1, 1),       # tiny map: several images per 128-row tile, ragged image count
    (5, 8, 64, 40, 3, 1, 2),       # Cout not a multiple of 32
])
def test_conv2d_implicit_gemm_matches_fp32_reference(pose, B, H, Cin, Cout, k, stride, dil):
    g = torch.Generator().manual_seed(B * 100 + H + k)
    x = (torch.randn(B, H, H, Cin, generator=g)).to(DEV).bfloat16()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(DEV).bfloat16()
    bias = torch.randn(Cout, generator=g).to(DEV)
    pad = (k - 1) // 2 * dil
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride, pad, dil)
    ref = F.silu(ref).permute(0, 2, 3, 1).contiguous()
    Ho = ref.shape[1]
    res = torch.randn(B, Ho, Ho, Cout, generator=g).to(DEV).bfloat16()
    out = torch.empty(B, Ho, Ho, Cout, device=DEV, dtype=torch.float32)
    e = _epi(pose, out, bias, 2, res, 0.5, 2.0)
    code = pose._lib.lib().pose_conv2d_bf16(x.data_ptr(), B, H, H, Cin, w.data_ptr(), Cout, k, k, stride, dil, pad,
                                            C.byref(e), pose._lib.stream_ptr())
    pose._lib.check(code, "pose_conv2d_bf16")
    want = ref * 0.5 + res.float() * 2.0
    assert torch.allclose(out, want, rtol=2e-3, atol=2e-3), (out - want).abs().max().item()


def test_gemm_epilogue_residual_and_concat(pose):
    g = torch.Generator().manual_seed(1)
    M, K, N = 384, 256, 96
    a = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).bfloat16()
    res = torch.randn(M, N, generator=g).to(DEV).bfloat16()
    cat = torch.full((M, 160), 3.0, device=DEV, dtype=torch.bfloat16)
    e = _epi(pose, cat, None, 0, res, 1.0, 0.25)
    e.C = cat.data_ptr() + 64 * 2          # column slice [64, 160)
    pose._lib.check(pose._lib.lib().pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e),
                                                      pose._lib.stream_ptr()), "gemm")
    want = a.float() @ w.float().t() + 0.25 * res.float()
    assert torch.allclose(cat[:, 64:].float(), want, rtol=2 ** -7, atol=2e-2)
    assert (cat[:, :64] == 3.0).all()


def test_dwconv_gates_pools_match_fp32_reference(pose):
    lib, sp = pose._lib.lib(), pose._lib.stream_ptr
    g = torch.Generator().manual_seed(2)
    for (B, H, Cc, stride) in [(2, 32, 64, 1), (3, 16, 384, 2), (2, 8, 3072, 1), (2, 17, 128, 2)]:
        x = torch.randn(B, H, H, Cc, generator=g).to(DEV).bfloat16()
        wd = (torch.randn(Cc, 1, 3, 3, generator=g) * 0.3).to(DEV)
        bias = torch.randn(Cc, generator=g).to(DEV)
        ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), wd, bias, stride, 1, groups=Cc)).permute(0, 2, 3, 1)
        Ho = ref.shape[1]
        y = torch.empty(B, Ho, Ho, Cc, device=DEV, dtype=torch.bfloat16)
        parts = lib.pose_dwconv3x3_pool_parts(H, H, stride)
        pool = torch.full((B, parts, Cc), float("nan"), device=DEV)    # written, not accumulated
        wk = wd.view(Cc, 9).t().contiguous()
        pose._lib.check(lib.pose_dwconv3x3_bf16(x.data_ptr(), B, H, H, Cc, wk.data_ptr(), bias.data_ptr(), stride, 2,
                                                y.data_ptr(), pool.data_ptr(), parts, sp()), "dw")
        assert torch.allclose(y.float(), ref, rtol=2 ** -7, atol=2e-2)
        assert torch.allclose(pool.sum(1), ref.sum((1, 2)), rtol=1e-3, atol=Ho * Ho * 2e-3)
        y2 = torch.empty_like(y)                                        # without the squeeze output
        pose._lib.check(lib.pose_dwconv3x3_bf16(x.data_ptr(), B, H, H, Cc, wk.data_ptr(), bias.data_ptr(), stride, 2,
                                                y2.data_ptr(), None, 0, sp()), "dw")
        assert torch.equal(y, y2)
        # pool_sum on its own
        pool2 = torch.full((B, 3, Cc), float("nan"), device=DEV)
        pose._lib.check(lib.pose_pool_sum_bf16(y.data_ptr(), B, Ho * Ho, Cc, pool2.data_ptr(), 3, sp()), "pool")
        assert torch.allclose(pool2.sum(1), y.float().sum((1, 2)), rtol=1e-4, atol=1e-2)
        # SE gate
        cr = max(2, Cc // 16)
        w1, w2 = (torch.randn(cr, Cc, generator=g) * 0.1).to(DEV), (torch.randn(Cc, cr, generator=g) * 0.1).to(DEV)
        gate = torch.empty(B, Cc, device=DEV)
        pose._lib.check(lib.pose_se_gate(pool2.data_ptr(), 3, 1.0 / (Ho * Ho), w1.data_ptr(), w2.data_ptr(), B, Cc, cr, 2,
                                         gate.data_ptr(), sp()), "se")
        mean = pool2.sum(1) / (Ho * Ho)
        want = torch.sigmoid(F.silu(mean @ w1.t()) @ w2.t())
        assert torch.allclose(gate, want, rtol=1e-4, atol=1e-5)
        # ECA gate (+ fused mean * gate)
        wk5 = torch.randn(5, generator=g).to(DEV)
        gate2 = torch.empty(B, Cc, device=DEV)
        feat = torch.empty(B, Cc, device=DEV, dtype=torch.bfloat16)
        pose._lib.check(lib.pose_eca_gate(pool2.data_ptr(), 3, 1.0 / (Ho * Ho), wk5.data_ptr(), 5, B, Cc, gate2.data_ptr(),
                                          feat.data_ptr(), sp()), "eca")
        m16 = torch.empty(B, Cc, device=DEV, dtype=torch.bfloat16)
        pose._lib.check(lib.pose_sums_to_bf16(pool2.data_ptr(), 3, B, Cc, 1.0 / (Ho * Ho), m16.data_ptr(), sp()), "s2b")
        assert torch.allclose(m16.float(), mean, rtol=2 ** -7, atol=1e-3)
        want2 = torch.sigmoid(F.conv1d(mean[:, None], wk5.view(1, 1, 5), padding=2))[:, 0]
        assert torch.allclose(gate2, want2, rtol=1e-4, atol=1e-5)
        assert torch.allclose(feat.float(), mean * want2, rtol=2 ** -7, atol=1e-3)
        # gate apply + broadcast add
        add = torch.randn(B, Cc, generator=g).to(DEV).bfloat16()
        z = torch.empty_like(y)
        pose._lib.check(lib.pose_channel_affine_bf16(y.data_ptr(), gate.data_ptr(), add.data_ptr(), B, Ho * Ho, Cc,
                                                     z.data_ptr(), sp()), "affine")
        assert torch.allclose(z.float(), y.float() * gate[:, None, None] + add.float()[:, None, None], rtol=2 ** -7, atol=2e-2)
        if Ho % 2 == 0:
            p = torch.empty(B, Ho // 2, Ho // 2, Cc, device=DEV, dtype=torch.bfloat16)
            pose._lib.check(lib.pose_avgpool2x2_bf16(y.data_ptr(), B, Ho, Ho, Cc, p.data_ptr(), sp()), "avgpool")
            wantp = F.avg_pool2d(y.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
            assert torch.allclose(p.float(), wantp, rtol=2 ** -7, atol=1e-2)
        # coordinate attention pieces
        P = torch.empty(B, 2 * Ho, Cc, device=DEV, dtype=torch.bfloat16)
        pose._lib.check(lib.pose_coord_pool_bf16(y.data_ptr(), B, Ho, Ho, Cc, P.data_ptr(), sp()), "coord_pool")
        yf = y.float()
        assert torch.allclose(P[:, :Ho].float(), yf.mean(2), rtol=2 ** -7, atol=1e-2)
        assert torch.allclose(P[:, Ho:].float(), yf.mean(1), rtol=2 ** -7, atol=1e-2)
        G = torch.rand(B, 2 * Ho, 2 * Cc, generator=g).to(DEV).bfloat16()
        o = torch.empty_like(y)
        pose._lib.check(lib.pose_coord_apply_bf16(y.data_ptr(), G.data_ptr(), B, Ho, Ho, Cc, o.data_ptr(), sp()), "coord_apply")
        wanto = yf * G[:, :Ho, None, :Cc].float() * G[:, None, Ho:, Cc:].float()
        assert torch.allclose(o.float(), wanto, rtol=2 ** -6, atol=2e-2)


def test_cnn_input_pack_matches_oracle(pose, oracle):
    rng = np.random.default_rng(0)
    B, S, J = 2, 64, 17
    img, dep = rng.random((B, 3, S, S), dtype=np.float32), rng.random((B, 1, S, S), dtype=np.float32)
    kp = rng.uniform(0.05, 0.95, (B, J, 2)).astype(np.float32)
    kp[0, 3] = -1
    out = torch.empty(B, S, S, 32, device=DEV, dtype=torch.bfloat16)
    ti, td, tk = (torch.from_numpy(a).to(DEV) for a in (img, dep, kp))   # keep the device copies alive
    pose._lib.check(pose._lib.lib().pose_cnn_input_pack(ti.data_ptr(), td.data_ptr(), tk.data_ptr(), B, S, J, 3.0,
                                                        out.data_ptr(), pose._lib.stream_ptr()), "pack")
    o = out.float().cpu().numpy()
    want = np.concatenate([img, dep, oracle.heatmap(kp, S, 3.0)], 1).transpose(0, 2, 3, 1)
    assert np.allclose(o[..., :21], want, rtol=2 ** -8, atol=1e-37)
    assert (o[..., 21:] == 0).all()


def _load_filled(pose, cfg, seed):
    from oracle import torch_models as tm
    m = pose.CNNPoseEstimation(cfg)
    sd = tm.fill_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    return m.to(DEV).eval(), sd


def test_cnn_small_matches_reference_golden(pose, golden):
    g = golden("cnn_small.npz")
    cfg = pose.ModelConfig("cnn", **json.loads(str(g["config"])))
    m, _ = _load_filled(pose, cfg, int(g["fill_seed"]))
    with torch.no_grad():
        out = m(torch.from_numpy(g["image"]).to(DEV), torch.from_numpy(g["depth"]).to(DEV), torch.from_numpy(g["kp"]).to(DEV))
    ref = torch.from_numpy(g["out"]).to(DEV)
    assert out.shape == ref.shape and out.dtype == torch.float32
    mpjpe = pose.utils.compute_mpjpe(out, ref).item()
    scale = ref.norm(dim=2).mean().item()
    print(f"small CNN: MPJPE vs reference {mpjpe:.4f} mm at joint magnitude {scale:.1f} mm")
    assert mpjpe < 0.5, mpjpe


@pytest.mark.parametrize("B", [2, 8])
def test_cnn_full_size_matches_fp32_oracle(pose, B):
    """BASELINE config 5 shape (256x256, 17 joints): bf16 tensor-core forward vs the fp32 oracle on the GPU."""
    from oracle import torch_models as tm
    cfg = pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256)
    m, sd = _load_filled(pose, cfg, 1)
    g = torch.Generator().manual_seed(B)
    img, dep = torch.rand(B, 3, 256, 256, generator=g).to(DEV), torch.rand(B, 1, 256, 256, generator=g).to(DEV)
    kp = (torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05).to(DEV)
    kp[0, 3] = -1.0
    with torch.no_grad():
        out = m(img, dep, kp)
        ref = tm.cnn_forward({k: v.to(DEV) for k, v in sd.items()}, cfg, img, dep, kp)
    mpjpe = pose.utils.compute_mpjpe(out, ref).item()
    scale = ref.norm(dim=2).mean().item()
    print(f"full CNN B={B}: MPJPE vs fp32 oracle {mpjpe:.4f} mm at joint magnitude {scale:.1f} mm")
    assert torch.isfinite(out).all()
    assert mpjpe < 0.5, (mpjpe, scale)
    # second call reuses the plan; parameters changed in place are picked up
    with torch.no_grad():
        out2 = m(img, dep, kp)
        assert torch.equal(out, out2)      # no atomics anywhere on the path: bit-reproducible
        m.pose_head.decoder[-1].bias.add_(10.0)
        out3 = m(img, dep, kp)
    assert torch.allclose(out3, out + 10.0, atol=1e-3)
    with pytest.raises(NotImplementedError), torch.no_grad():
        m.train()(img, dep, kp)            # batch statistics without a backward: eval() is the inference mode


def test_cnn_reference_default_500x500_forward_and_training_step(pose):
    """ModelConfig("cnn") defaults (src/model_config.py:59-66): 500 x 500 input, heat-map 500, AdaptiveAvgPool2d(8) from a
    32 x 32 map (cnn.py:602).  Maps of 250 / 125 / 63 / 32 pixels: ragged implicit-GEMM patches, odd depthwise sizes, the
    general adaptive pooling.  Eval forward within the 0.5 mm bar of the fp32 oracle; one training step: loss, every
    parameter gradient against fp32 autograd (bf16 bars at batch 2: norm within 25 %, mean relative error no larger than
    torch.autocast(bfloat16)'s on the same model)."""
    from oracle import torch_models as tm
    cfg = pose.ModelConfig("cnn", regression_dropout=0.0)
    assert tuple(cfg.image_size) == (500, 500) and int(cfg.heatmap_size) == 500 and int(cfg.global_pool_size) == 8
    m, sd = _load_filled(pose, cfg, 3)
    sd = {k: v.to(DEV) for k, v in sd.items()}
    B = 2
    g = torch.Generator().manual_seed(77)
    img, dep = torch.rand(B, 3, 500, 500, generator=g).to(DEV), torch.rand(B, 1, 500, 500, generator=g).to(DEV)
    kp = (torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05).to(DEV)
    gt = (torch.randn(B, 17, 3, generator=g) * 300).to(DEV)
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            out = m(img, dep, kp)
            ref = tm.cnn_forward(sd, cfg, img, dep, kp)
        mpjpe = pose.utils.compute_mpjpe(out, ref).item()
        print(f"500x500 CNN eval: MPJPE vs fp32 oracle {mpjpe:.4f} mm at joint magnitude {ref.norm(dim=2).mean().item():.1f} mm")
        assert torch.isfinite(out).all() and mpjpe < 0.5, mpjpe
        # training step
        m.train()
        crit = pose.ComprehensivePoseLoss()
        pred = m(img, dep, kp)
        total, _ = crit(pred, gt)
        total.backward()
        names = [n for n, _ in m.named_parameters()]
        sdg = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
        po, _ = tm.cnn_forward(sdg, cfg, img, dep, kp, train=True, return_stats=True)
        lref = tm.composite_loss(po, gt)
        lref.backward()
        # the same step under torch.autocast(bfloat16): the yardstick for what bf16 storage costs on this model
        sda = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pa, _ = tm.cnn_forward(sda, cfg, img, dep, kp, train=True, return_stats=True)
        tm.composite_loss(pa.float(), gt).backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert abs(total.item() - lref.item()) < 2e-2 * abs(lref.item()), (total.item(), lref.item())
    gmax = max(sdg[n].grad.norm().item() for n in names)
    errs, worst, norm_dev, auto = [], (0.0, None), (0.0, None), []
    for n, p in m.named_parameters():
        r = sdg[n].grad.double()
        rn = r.norm().item()
        if rn < 1e-5 * gmax:
            continue                                     # analytically-zero gradients hold rounding noise
        an = p.grad.double().norm().item()
        norm_dev = max(norm_dev, (abs(an - rn) / rn, n))
        e = (p.grad.double() - r).norm().item() / rn
        errs.append(e)
        auto.append((sda[n].grad.double() - r).norm().item() / rn)
        worst = max(worst, (e, n))
    mean_ours, mean_auto = sum(errs) / len(errs), sum(auto) / len(auto)
    print(f"500x500 CNN train: mean relative gradient error {mean_ours:.4f} (autocast {mean_auto:.4f}), worst {worst}, "
          f"worst norm deviation {norm_dev}")
    assert norm_dev[0] < 0.25, norm_dev
    assert mean_ours < mean_auto + 0.01, (mean_ours, mean_auto, worst)
