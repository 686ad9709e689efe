"""GPU: TransformerPoseEstimation -- the non-GEMM kernels against plain PyTorch fp32 references of the same ops, the
eval-mode forward against the golden frozen from the live reference (0.5 mm MPJPE, BASELINE.json:north_star), and one
training step (loss, every parameter's gradient, fused AdamW) against fp32 autograd over the oracle restatement, which
gen_golden.py pins to the live reference's gradients."""
import importlib
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib(pose):
    return pose._lib.lib(), pose._lib.stream_ptr, pose._lib.check


@pytest.mark.parametrize("D,M", [(768, 1000), (256, 77), (1024, 64)])
def test_layernorm_forward_backward(pose, D, M):
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(D + M)
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).to(DEV).bfloat16()
    gamma = (torch.rand(D, generator=g) + 0.5).to(DEV)
    beta = torch.randn(D, generator=g).to(DEV)
    dy = torch.randn(M, D, generator=g).to(DEV).bfloat16()
    dres = torch.randn(M, D, generator=g).to(DEV).bfloat16()
    y = torch.empty_like(x)
    check(lib.pose_layernorm_bf16(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, M, M, M, 0, M, 0, D,
                                  y.data_ptr(), sp()), "ln")
    xr = x.float().requires_grad_()
    gr, br = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    yr = F.layer_norm(xr, (D,), gr, br, 1e-5)
    assert torch.allclose(y.float(), yr, rtol=2 ** -7, atol=2e-2)
    yr.backward(dy.float())
    dx = torch.empty_like(x)
    dg, db = torch.full((D,), 1.0, device=DEV), torch.zeros(D, device=DEV)
    check(lib.pose_layernorm_bwd_bf16(x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), 1e-5, M, M, M, 0, M, 0, D,
                                      dres.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), sp()), "ln_bwd")
    assert torch.allclose(dx.float(), xr.grad + dres.float(), rtol=2e-2, atol=3e-2)
    assert torch.allclose(dg - 1.0, gr.grad, rtol=1e-3, atol=1e-2 * M ** 0.5)
    assert torch.allclose(db, br.grad, rtol=1e-3, atol=1e-2 * M ** 0.5)


def test_layernorm_token_subrange_addressing(pose):
    """drop the class token of every sample (transformers.py:336-346) / normalise only token 0 (:371)."""
    lib, sp, check = _lib(pose)
    B, T, D = 3, 17, 768
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, D, generator=g).to(DEV).bfloat16()
    gamma, beta = (torch.rand(D, generator=g) + 0.5).to(DEV), torch.randn(D, generator=g).to(DEV)
    y = torch.empty(B, T - 1, D, device=DEV, dtype=torch.bfloat16)
    check(lib.pose_layernorm_bf16(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, B * (T - 1), T - 1, T, 1, T - 1,
                                  0, D, y.data_ptr(), sp()), "ln")
    want = F.layer_norm(x[:, 1:].float(), (D,), gamma, beta, 1e-6)
    assert torch.allclose(y.float(), want, rtol=2 ** -7, atol=2e-2)
    y0 = torch.empty(B, D, device=DEV, dtype=torch.bfloat16)
    check(lib.pose_layernorm_bf16(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-5, B, 1, T, 0, 1, 0, D,
                                  y0.data_ptr(), sp()), "ln")
    assert torch.allclose(y0.float(), F.layer_norm(x[:, 0].float(), (D,), gamma, beta, 1e-5), rtol=2 ** -7, atol=2e-2)


@pytest.mark.parametrize("hd,heads,Nq,Nk", [(64, 12, 257, 257), (48, 16, 273, 273), (48, 16, 256, 16), (48, 16, 16, 256),
                                            (64, 2, 70, 33),
                                            # the reference's default 512 x 512 input: 1025 backbone tokens, 1024 + 64 + 1 final
                                            (64, 12, 1025, 1025), (48, 16, 1089, 1089), (48, 4, 1024, 64), (48, 4, 64, 1024)])
def test_attention_forward_backward(pose, hd, heads, Nq, Nk):
    lib, sp, check = _lib(pose)
    B, E = 2, hd * heads
    g = torch.Generator().manual_seed(Nq * 7 + Nk)
    q = torch.randn(B, Nq, E, generator=g).to(DEV).bfloat16()
    kv = torch.randn(B, Nk, 2 * E, generator=g).to(DEV).bfloat16()        # packed k | v, pitch 2E
    do = torch.randn(B, Nq, E, generator=g).to(DEV).bfloat16()
    o = torch.empty(B, Nq, E, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, Nq, device=DEV)
    scale = 1.0 / math.sqrt(hd)
    k_ptr, v_ptr = kv.data_ptr(), kv.data_ptr() + 2 * E
    check(lib.pose_attention_bf16(q.data_ptr(), k_ptr, v_ptr, o.data_ptr(), B, heads, Nq, Nk, hd, E, 2 * E, 2 * E, E,
                                  Nq * E, Nk * 2 * E, Nk * 2 * E, Nq * E, scale, lse.data_ptr(), 0.0, 0, sp()), "attn")
    qr = q.float().requires_grad_()
    kvr = kv.float().requires_grad_()
    qh = qr.view(B, Nq, heads, hd).transpose(1, 2)
    kh = kvr[..., :E].reshape(B, Nk, heads, hd).transpose(1, 2)
    vh = kvr[..., E:].reshape(B, Nk, heads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    orf = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Nq, E)
    assert torch.allclose(o.float(), orf, rtol=2e-2, atol=2e-2), (o.float() - orf).abs().max().item()
    assert torch.allclose(lse, torch.logsumexp(s, -1), rtol=1e-3, atol=1e-2)
    orf.backward(do.float())
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    dws = torch.empty(B, heads, Nq, device=DEV)
    check(lib.pose_attention_bwd_bf16(q.data_ptr(), k_ptr, v_ptr, o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                      dq.data_ptr(), dkv.data_ptr(), dkv.data_ptr() + 2 * E, dws.data_ptr(), B, heads, Nq,
                                      Nk, hd, E, 2 * E, 2 * E, E, E, E, 2 * E, 2 * E, Nq * E, Nk * 2 * E, Nk * 2 * E,
                                      Nq * E, Nq * E, Nq * E, Nk * 2 * E, Nk * 2 * E, scale, 0.0, 0, sp()), "attn_bwd")
    tol = dict(rtol=3e-2, atol=3e-2)
    assert torch.allclose(dq.float(), qr.grad, **tol), (dq.float() - qr.grad).abs().max().item()
    assert torch.allclose(dkv.float(), kvr.grad, **tol), (dkv.float() - kvr.grad).abs().max().item()


def test_attention_class_token_outside_the_tiles(pose, monkeypatch):
    """257 = 2 x 128 + 1 tokens: the tensor-core tiles run on tokens 1..256 and the class token is handled on the CUDA cores
    (AttnTail in csrc/attention_tc.cu); the mode needs a large batch by default, POSE_ATTN_TAIL=2 forces it."""
    monkeypatch.setenv("POSE_ATTN_TAIL", "2")
    test_attention_forward_backward(pose, 64, 12, 257, 257)
    monkeypatch.setenv("POSE_ATTN_TAIL", "0")
    test_attention_forward_backward(pose, 64, 12, 257, 257)


@pytest.mark.parametrize("hd,heads,Nq,Nk", [(48, 16, 273, 273), (64, 12, 257, 257), (48, 16, 16, 256), (64, 3, 1025, 1025)])
def test_attention_weight_dropout_forward_backward(pose, hd, heads, Nq, Nk):
    """nn.MultiheadAttention(dropout=p): the mask is a counter-based hash of (seed, element); pose_dropout_bf16 over a
    tensor of ones with the same seed exposes it, so the fused kernels can be checked against plain PyTorch."""
    lib, sp, check = _lib(pose)
    B, E, p, seed = 2, hd * heads, 0.1, 1234567
    g = torch.Generator().manual_seed(Nq + Nk)
    q = torch.randn(B, Nq, E, generator=g).to(DEV).bfloat16()
    kv = torch.randn(B, Nk, 2 * E, generator=g).to(DEV).bfloat16()
    do = torch.randn(B, Nq, E, generator=g).to(DEV).bfloat16()
    ones = torch.ones(B, heads, Nq, Nk, device=DEV, dtype=torch.bfloat16)
    mask = torch.empty_like(ones)
    check(lib.pose_dropout_bf16(ones.data_ptr(), ones.numel(), p, seed, mask.data_ptr(), sp()), "dropout")
    keep = (mask != 0).float().mean().item()
    assert abs(keep - (1 - p)) < 0.01 and abs(mask.float().max().item() - 1 / (1 - p)) < 0.01
    o = torch.empty(B, Nq, E, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, Nq, device=DEV)
    scale = 1.0 / math.sqrt(hd)
    k_ptr, v_ptr = kv.data_ptr(), kv.data_ptr() + 2 * E
    check(lib.pose_attention_bf16(q.data_ptr(), k_ptr, v_ptr, o.data_ptr(), B, heads, Nq, Nk, hd, E, 2 * E, 2 * E, E,
                                  Nq * E, Nk * 2 * E, Nk * 2 * E, Nq * E, scale, lse.data_ptr(), p, seed, sp()), "attn")
    qr, kvr = q.float().requires_grad_(), kv.float().requires_grad_()
    qh = qr.view(B, Nq, heads, hd).transpose(1, 2)
    kh = kvr[..., :E].reshape(B, Nk, heads, hd).transpose(1, 2)
    vh = kvr[..., E:].reshape(B, Nk, heads, hd).transpose(1, 2)
    pr = torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1) * mask.float()
    orf = (pr @ vh).transpose(1, 2).reshape(B, Nq, E)
    assert torch.allclose(o.float(), orf, rtol=2e-2, atol=2e-2), (o.float() - orf).abs().max().item()
    orf.backward(do.float())
    dq, dkv, dws = torch.empty_like(q), torch.empty_like(kv), torch.empty(B, heads, Nq, device=DEV)
    check(lib.pose_attention_bwd_bf16(q.data_ptr(), k_ptr, v_ptr, o.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                      dq.data_ptr(), dkv.data_ptr(), dkv.data_ptr() + 2 * E, dws.data_ptr(), B, heads, Nq,
                                      Nk, hd, E, 2 * E, 2 * E, E, E, E, 2 * E, 2 * E, Nq * E, Nk * 2 * E, Nk * 2 * E,
                                      Nq * E, Nq * E, Nq * E, Nk * 2 * E, Nk * 2 * E, scale, p, seed, sp()), "attn_bwd")
    tol = dict(rtol=3e-2, atol=3e-2)
    assert torch.allclose(dq.float(), qr.grad, **tol), (dq.float() - qr.grad).abs().max().item()
    assert torch.allclose(dkv.float(), kvr.grad, **tol), (dkv.float() - kvr.grad).abs().max().item()


def test_gemm_epilogue_dropout_matches_the_elementwise_mask(pose):
    import ctypes as C
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(5)
    M, K, N, p, seed = 300, 256, 768, 0.25, 99
    a = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV).bfloat16()
    res = torch.randn(M, N, generator=g).to(DEV).bfloat16()

    def run(drop_p):
        out = torch.empty(M, N, device=DEV, dtype=torch.float32)
        e = pose._lib.PoseGemmEpilogue()
        e.C, e.ldc, e.act, e.out_dtype, e.out_scale = out.data_ptr(), N, 3, 0, 1.0
        e.residual, e.ldr, e.res_scale = res.data_ptr(), N, 1.0
        e.drop_p, e.drop_seed = drop_p, seed
        check(lib.pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e), sp()), "gemm")
        return out
    plain, dropped = run(0.0), run(p)
    ones = torch.ones(M, N, device=DEV, dtype=torch.bfloat16)
    mask = torch.empty_like(ones)
    check(lib.pose_dropout_bf16(ones.data_ptr(), ones.numel(), p, seed, mask.data_ptr(), sp()), "dropout")
    want = (plain - res.float()) * (mask != 0).float() / (1 - p) + res.float()       # x + drop(gelu(a w^T))
    assert torch.allclose(dropped, want, rtol=1e-4, atol=1e-4)


def test_sums_slices_and_padded_cast(pose):
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(8)
    M, N, ld = 1000, 51, 56
    x = torch.zeros(M, ld)
    x[:, :N] = torch.randn(M, N, generator=g)
    x = x.to(DEV).bfloat16()
    out = torch.full((N,), 2.0, device=DEV)
    check(lib.pose_colsum_bf16(x.data_ptr(), M, N, ld, out.data_ptr(), sp()), "colsum")
    assert torch.allclose(out, 2.0 + x[:, :N].float().sum(0), rtol=1e-4, atol=1e-3)
    B, T, D = 5, 19, 768
    t = torch.randn(B, T, D, generator=g).to(DEV).bfloat16()
    acc = torch.ones(T - 2, D, device=DEV)
    check(lib.pose_batch_rowsum_bf16(t.data_ptr(), B, T, 2, T - 2, D, acc.data_ptr(), sp()), "rowsum")
    assert torch.allclose(acc, 1.0 + t[:, 2:].float().sum(0), rtol=1e-4, atol=1e-3)
    sl = torch.empty(B, 4, D, device=DEV, dtype=torch.bfloat16)
    check(lib.pose_token_slice_bf16(t.data_ptr(), B, T, 3, 4, D, sl.data_ptr(), sp()), "slice")
    assert torch.equal(sl, t[:, 3:7])
    f = torch.randn(7, N, generator=g).to(DEV)
    o = torch.full((7, ld), 9.0, device=DEV, dtype=torch.bfloat16)
    check(lib.pose_cast_f32_bf16_2d(f.data_ptr(), N, 7, N, o.data_ptr(), ld, sp()), "cast2d")
    assert torch.equal(o[:, :N], f.bfloat16()) and (o[:, N:] == 0).all()


def test_fused_adamw_matches_torch(pose):
    g = torch.Generator().manual_seed(12)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in [(300, 17), (51,), (64, 64, 3)]]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = pose.AdamW(ps, lr=1e-2, weight_decay=0.05)
    ropt = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.05)
    flat = pose.params.FlatParams.of(ps)
    for step in range(3):
        for p, r in zip(ps, ref):
            gr = torch.randn(p.shape, generator=g).to(DEV)
            p.grad.copy_(gr)
            r.grad = gr.clone()
        opt.step()
        ropt.step()
        for p, r in zip(ps, ref):
            assert torch.allclose(p, r, rtol=1e-5, atol=1e-6)
            assert torch.equal(flat.w16(p).view(p.shape), p.detach().bfloat16())   # shadow refreshed by the kernel
            assert (p.grad == 0).all()                                             # and the gradient cleared


def test_optimizer_state_round_trip_with_torch_adamw(pose, tmp_path):
    """optimizer_state_dict of the reference's checkpoints (src/train.py:300-306, main.py:130-134): the fused AdamW saves
    torch.optim.AdamW's per-parameter layout and loads it back.  (i) save -> load into a fresh fused optimizer -> the
    continued run is bit-identical to the uninterrupted one; (ii) the saved dict loads into torch.optim.AdamW and that
    continues the same trajectory; (iii) a state dict written by torch.optim.AdamW resumes the fused optimizer."""
    g = torch.Generator().manual_seed(21)
    shapes = [(300, 17), (51,), (64, 64, 3)]
    init = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(6)]

    def fresh(kind):
        ps = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
        opt = (pose.AdamW if kind == "fused" else torch.optim.AdamW)(ps, lr=1e-2, weight_decay=0.05)
        return ps, opt

    def run(ps, opt, steps):
        for k in steps:
            for p, gr in zip(ps, grads[k]):
                if p.grad is None:
                    p.grad = gr.clone().to(DEV)
                else:
                    p.grad.copy_(gr.to(DEV))
            opt.step()

    # uninterrupted reference trajectories
    ps_a, opt_a = fresh("fused")
    pose.params.FlatParams.of(ps_a)
    run(ps_a, opt_a, range(6))
    # (i) fused -> file -> fused
    ps_b, opt_b = fresh("fused")
    pose.params.FlatParams.of(ps_b)
    run(ps_b, opt_b, range(3))
    sd = opt_b.state_dict()
    assert set(sd) == {"state", "param_groups"} and sorted(sd["state"]) == [0, 1, 2]
    assert all(set(v) == {"step", "exp_avg", "exp_avg_sq"} and float(v["step"]) == 3.0 for v in sd["state"].values())
    assert [tuple(sd["state"][i]["exp_avg"].shape) for i in range(3)] == shapes
    path = tmp_path / "opt.pth"
    torch.save({"optimizer_state_dict": sd, "params": [p.detach().cpu() for p in ps_b]}, path)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    ps_c = [torch.nn.Parameter(t.clone().to(DEV)) for t in ck["params"]]
    opt_c = pose.AdamW(ps_c, lr=1e-2, weight_decay=0.05)
    pose.params.FlatParams.of(ps_c)
    opt_c.load_state_dict(ck["optimizer_state_dict"])
    run(ps_c, opt_c, range(3, 6))
    for a, c in zip(ps_a, ps_c):
        assert torch.equal(a, c)
    # (ii) fused -> torch.optim.AdamW
    ps_t = [torch.nn.Parameter(t.clone().to(DEV)) for t in ck["params"]]
    opt_t = torch.optim.AdamW(ps_t, lr=1e-2, weight_decay=0.05)
    opt_t.load_state_dict(ck["optimizer_state_dict"])
    run(ps_t, opt_t, range(3, 6))
    for a, t in zip(ps_a, ps_t):
        assert torch.allclose(a, t, rtol=1e-5, atol=1e-6)
    # (iii) torch.optim.AdamW -> fused
    ps_r, opt_r = fresh("torch")
    run(ps_r, opt_r, range(3))
    ps_f = [torch.nn.Parameter(p.detach().clone()) for p in ps_r]
    opt_f = pose.AdamW(ps_f, lr=1e-2, weight_decay=0.05)
    pose.params.FlatParams.of(ps_f)
    opt_f.load_state_dict(opt_r.state_dict())
    run(ps_f, opt_f, range(3, 6))
    for a, f in zip(ps_a, ps_f):
        assert torch.allclose(a, f, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
def _model(pose, **kw):
    cfg = pose.ModelConfig("transformer", image_size=(256, 256), vit_pretrained=False, **kw)
    from oracle import torch_models as tm
    m = pose.TransformerPoseEstimation(cfg)
    sd = tm.fill_vit_state_dict(m.state_dict(), seed=7)
    m.load_state_dict(sd)
    return m.to(DEV), {k: v.to(DEV) for k, v in sd.items()}, tm


def _inputs(golden):
    gd = golden("vit_256.npz")
    g = torch.Generator().manual_seed(int(gd["image_seed"]))
    img = torch.rand(2, 3, 256, 256, generator=g)
    dep = torch.rand(2, 1, 256, 256, generator=g)
    kp = torch.from_numpy(gd["kp"])
    return gd, img.to(DEV), dep.to(DEV), kp.to(DEV)


def test_state_dict_layout_matches_reference(pose):
    m = pose.TransformerPoseEstimation(pose.ModelConfig("transformer", image_size=(256, 256), vit_pretrained=False))
    with open(os.path.join(os.path.dirname(__file__), "golden", "vit_state_dict_layout.json")) as f:
        layout = json.load(f)
    sd = m.state_dict()
    assert set(sd) == set(layout)
    assert all(list(sd[k].shape) == layout[k] for k in sd)


def test_eval_forward_matches_reference_golden(pose, golden):
    gd, img, dep, kp = _inputs(golden)
    m, sd, tm = _model(pose)
    m.eval()
    with torch.no_grad():
        out = m(img, dep, kp)
        ref = tm.vit_forward(sd, m.config, img, dep, kp)
    want = torch.from_numpy(gd["out"]).to(DEV)
    assert (ref - want).abs().max() < 2e-3 * want.abs().max()           # oracle on this GPU == live reference
    mpjpe = pose.utils.compute_mpjpe(out, want).item()
    assert mpjpe < 0.5, f"MPJPE vs the reference {mpjpe:.3f} mm"       # BASELINE.json:north_star


def test_training_step_gradients_match_fp32_autograd(pose, golden):
    gd, img, dep, kp = _inputs(golden)
    m, sd, tm = _model(pose, transformer_dropout_rate=0.0, transformer_attention_dropout_rate=0.0, regression_dropout=0.0)
    m.train()
    gt = torch.from_numpy(gd["gt"]).to(DEV)
    crit = pose.ComprehensivePoseLoss()
    pred = m(img, dep, kp)
    total, comps = crit(pred, gt)
    total.backward()
    assert abs(total.item() - float(gd["train_loss"])) < 2e-2 * float(gd["train_loss"])
    # fp32 autograd over the oracle restatement (pinned to the live reference by oracle/gen_golden.py)
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() and "grid" not in k else v) for k, v in sd.items()}
    po = tm.vit_forward(sdg, m.config, img, dep, kp)
    d = po - gt
    iu = torch.triu_indices(17, 17, 1, device=DEV)
    pd = lambda t: torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]   # noqa: E731
    lo = (d ** 2).mean() + d.abs().mean() + 100.0 * (pd(po) - pd(gt)).abs().mean() + d[:, 0].abs().mean()
    lo.backward()
    names = list(gd["grad_names"])
    gn_ref = dict(zip(names, gd["grad_norms"]))
    worst = []
    # gradients six orders of magnitude below the largest one (the LayerNorm in front of the 16-query cross attention
    # over 256 near-uniform keys) are below bf16 resolution: errors are measured against max(|ref|, 1e-6 * largest)
    floor = 1e-6 * float(max(gd["grad_norms"]))
    for n, p in m.named_parameters():
        g, r = p.grad.double(), sdg[n].grad.double()
        rn = r.norm().item()
        assert abs(rn - gn_ref[n]) <= 5e-3 * gn_ref[n] + 1e-6, (n, rn, gn_ref[n])      # oracle == live reference
        rel = (g - r).norm().item() / (rn + floor)
        worst.append((rel, n))
    worst.sort(reverse=True)
    assert worst[0][0] < 0.08, worst[:8]        # bf16 activations / weights end to end vs fp32
    assert sum(r for r, _ in worst) / len(worst) < 0.04, worst[:8]


def test_default_512x512_configuration_forward_and_gradients(pose):
    """ModelConfig("transformer") defaults to 512 x 512 (model_config.py): 1025 backbone tokens, 1024 + 64 + 1 tokens in
    the final encoder -- the streaming attention kernels.  Eval forward and every gradient against fp32 autograd over
    the oracle restatement (pinned to the live reference at 256 x 256 by the tests above)."""
    from oracle import torch_models as tm
    cfg = pose.ModelConfig("transformer", vit_pretrained=False, transformer_dropout_rate=0.0,
                           transformer_attention_dropout_rate=0.0, regression_dropout=0.0)
    assert tuple(cfg.image_size) == (512, 512)
    m = pose.TransformerPoseEstimation(cfg)
    sd = tm.fill_vit_state_dict(m.state_dict(), seed=11)
    m.load_state_dict(sd)
    m = m.to(DEV)
    sd = {k: v.to(DEV) for k, v in sd.items()}
    g = torch.Generator().manual_seed(5)
    img, dep = torch.rand(1, 3, 512, 512, generator=g).to(DEV), torch.rand(1, 1, 512, 512, generator=g).to(DEV)
    kp = (torch.rand(1, 17, 2, generator=g) * 0.9 + 0.05).to(DEV)
    gt = (torch.randn(1, 17, 3, generator=g) * 300).to(DEV)
    m.eval()
    with torch.no_grad():
        out = m(img, dep, kp)
        ref = tm.vit_forward(sd, cfg, img, dep, kp)
    mpjpe = pose.utils.compute_mpjpe(out, ref).item()
    assert mpjpe < 0.5, f"MPJPE vs the fp32 oracle {mpjpe:.3f} mm"
    m.train()
    total, _ = pose.ComprehensivePoseLoss()(m(img, dep, kp), gt)
    total.backward()
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() and "grid" not in k else v) for k, v in sd.items()}
    po = tm.vit_forward(sdg, cfg, img, dep, kp)
    d = po - gt
    iu = torch.triu_indices(17, 17, 1, device=DEV)
    pd = lambda t: torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]   # noqa: E731
    lo = (d ** 2).mean() + d.abs().mean() + 100.0 * (pd(po) - pd(gt)).abs().mean() + d[:, 0].abs().mean()
    lo.backward()
    assert abs(total.item() - lo.item()) < 2e-2 * lo.item()
    norms = {n: sdg[n].grad.double().norm().item() for n, _ in m.named_parameters()}
    floor = 1e-6 * max(norms.values())
    worst = sorted(((p.grad.double() - sdg[n].grad.double()).norm().item() / (norms[n] + floor), n) for n, p in m.named_parameters())
    assert worst[-1][0] < 0.08, worst[-8:]
    assert sum(r for r, _ in worst) / len(worst) < 0.04, worst[-8:]


def test_training_with_the_reference_dropout_rates(pose, golden):
    """Default config (transformer dropout 0.1, attention dropout 0.1, head dropout 0.25): the step runs, eval ignores
    dropout, training forwards differ from step to step (fresh masks) and the loss goes down."""
    gd, img, dep, kp = _inputs(golden)
    m, sd, tm = _model(pose)
    gt = torch.from_numpy(gd["gt"]).to(DEV)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(img, dep, kp), m(img, dep, kp)
    assert torch.equal(e1, e2)
    m.train()
    train = __import__("importlib").import_module("3dhumanposeestimation_b200.train")
    tr = train.Trainer(m, pose.ComprehensivePoseLoss(), lr=1e-4)
    plan = m.plan(2, img.device)
    a = plan.forward(img, dep, kp, save=True).clone()
    b = plan.forward(img, dep, kp, save=True).clone()
    assert not torch.equal(a, b) and torch.isfinite(a).all()
    losses = [tr.step(img, dep, kp, gt)[4].item() for _ in range(6)]
    assert all(np.isfinite(losses)) and min(losses[3:]) < losses[0], losses
    for p_ in m.parameters():
        assert torch.isfinite(p_).all()


def test_fused_optimizer_step_changes_the_next_forward(pose, golden):
    gd, img, dep, kp = _inputs(golden)
    m, sd, tm = _model(pose, transformer_dropout_rate=0.0, transformer_attention_dropout_rate=0.0, regression_dropout=0.0)
    m.train()
    gt = torch.from_numpy(gd["gt"]).to(DEV)
    crit = pose.ComprehensivePoseLoss()
    opt = pose.AdamW(m.parameters(), lr=1e-4, weight_decay=0.01)
    losses = []
    for _ in range(4):
        total, _ = crit(m(img, dep, kp), gt)
        total.backward()
        opt.step()
        opt.zero_grad()
        losses.append(total.item())
    assert losses[-1] < losses[0], losses


def test_patchify_and_fused_heatmap_patch_operand(pose, oracle):
    """The patch-embedding operands (transformers.py:41-46): (i) image + depth planes -> [B * patches, 4 * P * P] bf16, bit-equal
    to torch's unfold of the concatenated planes; (ii) the Gaussians rendered straight into the heat-map stream's operand
    equal the bf16 rounding of the reference's fp32 heat-maps (the C oracle's values, <= 2 ulp from the planes kernel),
    patch by patch -- the fp32 planes never exist."""
    lib, sp, check = _lib(pose)
    g = torch.Generator().manual_seed(31)
    B, H, P = 3, 64, 16
    img, dep = torch.rand(B, 3, H, H, generator=g).to(DEV), torch.rand(B, 1, H, H, generator=g).to(DEV)
    out = torch.empty(B * (H // P) ** 2, 4 * P * P, device=DEV, dtype=torch.bfloat16)
    check(lib.pose_patchify_bf16(img.data_ptr(), 3, dep.data_ptr(), 1, B, H, H, P, out.data_ptr(), sp()), "patchify")
    x = torch.cat([img, dep], 1)
    want = torch.nn.functional.unfold(x, P, stride=P).transpose(1, 2).reshape(B * (H // P) ** 2, 4 * P * P).bfloat16()
    assert torch.equal(out, want)
    # heat-map stream: hs 64, sigma 2, 16 x 16 patches (the reference's ViT configuration) and an 8-pixel patch variant
    J = 17
    kp = (torch.rand(B, J, 2, generator=g) * 0.9 + 0.05)
    kp[0, 3] = -1.0                                   # invalid key-point: an all-zero plane
    for hs, sigma, hp in ((64, 2.0, 16), (32, 1.5, 8)):
        planes = torch.empty(B, J, hs, hs, device=DEV)
        check(lib.pose_heatmap_render(kp.to(DEV).data_ptr(), B, J, hs, sigma, planes.data_ptr(), 0, 0, 0, 0, sp()), "render")
        fused = torch.empty(B * (hs // hp) ** 2, J * hp * hp, device=DEV, dtype=torch.bfloat16)
        check(lib.pose_heatmap_patchify_bf16(kp.to(DEV).data_ptr(), B, J, hs, sigma, hp, fused.data_ptr(), sp()), "hm patchify")
        want = torch.nn.functional.unfold(planes, hp, stride=hp).transpose(1, 2).reshape(fused.shape).bfloat16()
        assert torch.equal(fused, want)               # same arithmetic as the planes kernel, rounded once to bf16
        ref = torch.from_numpy(oracle.heatmap(kp.numpy(), hs, sigma)).to(DEV)
        refp = torch.nn.functional.unfold(ref, hp, stride=hp).transpose(1, 2).reshape(fused.shape)
        assert (fused.float() - refp).abs().max().item() <= 2 ** -8       # bf16 rounding of values in [0, 1]
        assert fused.float().sum().item() > 0


def test_vit_freeze_backbone_trains_everything_else(pose):
    """vit_freeze_backbone (transformers.py:226-236): the backbone keeps its weights -- except the patch embedding adapted to
    RGB-D -- while every other parameter receives the same gradient and update as in the unfrozen model."""
    train = importlib.import_module("3dhumanposeestimation_b200.train")
    kw = dict(image_size=(256, 256), vit_pretrained=False, transformer_dropout_rate=0.0, transformer_attention_dropout_rate=0.0,
              regression_dropout=0.0)
    torch.manual_seed(11)
    mf = pose.TransformerPoseEstimation(pose.ModelConfig("transformer", vit_freeze_backbone=True, **kw)).to(DEV).train()
    mu = pose.TransformerPoseEstimation(pose.ModelConfig("transformer", **kw)).to(DEV).train()
    mu.load_state_dict(mf.state_dict())
    frozen = {n for n, p in mf.named_parameters() if not p.requires_grad}
    assert frozen and all(n.startswith("vit_backbone.") and not n.startswith("vit_backbone.patch_embed.proj") for n in frozen)
    assert mf.vit_backbone.patch_embed.proj.weight.requires_grad
    B = 2
    g = torch.Generator().manual_seed(5)
    img, dep = torch.rand(B, 3, 256, 256, generator=g).to(DEV), torch.rand(B, 1, 256, 256, generator=g).to(DEV)
    kp = (torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05).to(DEV)
    gt = (torch.randn(B, 17, 3, generator=g) * 300).to(DEV)
    before = {n: p.detach().clone() for n, p in mf.named_parameters()}
    tf = train.Trainer(mf, pose.ComprehensivePoseLoss(), lr=1e-3, weight_decay=0.01, graph=False)
    tu = train.Trainer(mu, pose.ComprehensivePoseLoss(), lr=1e-3, weight_decay=0.01, graph=False)
    lf, lu = tf.step(img, dep, kp, gt)[4].item(), tu.step(img, dep, kp, gt)[4].item()
    assert abs(lf - lu) <= 1e-5 * abs(lu)
    pu = dict(mu.named_parameters())
    for n, p in mf.named_parameters():
        if n in frozen:
            assert torch.equal(p, before[n]), n                   # no update, no weight decay
        else:
            assert not torch.equal(p, before[n]), n
            # same update as the unfrozen model (split-K atomics reorder fp32 sums: AdamW's first step is ~lr * sign(g))
            assert (p - pu[n]).abs().max().item() <= 2.1e-3, n
            assert ((p - pu[n]).abs() > 1e-4).float().mean().item() < 0.02, n
    sd = tf.opt.state_dict()
    assert len(sd["state"]) == sum(1 for p in mf.parameters() if p.requires_grad)
