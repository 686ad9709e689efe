"""GPU: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs, against the
goldens frozen from the live reference, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Tolerances are the ones BASELINE.json:north_star states."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


# ------------------------------------------------------------------------------------------------ loss
def _loss_case(rng, B, J=17):
    gt = rng.normal(0, 300, (B, J, 3)).astype(np.float32)
    pred = (gt + rng.normal(0, 50, (B, J, 3))).astype(np.float32)
    if J > 5:
        pred[0, 3] = pred[0, 4]  # coincident joints
        pred[0, 2] = gt[0, 2]    # zero difference
    return pred, gt


@pytest.mark.parametrize("B", [1, 7, 8, 256, 5000])
def test_loss_matches_oracle(pose, oracle, B):
    rng = np.random.default_rng(B)
    pred, gt = _loss_case(rng, B)
    w = (1.0, 1.0, 100.0, 1.0)
    ref5, refg = oracle.pose_loss(pred, gt, w)
    p = _t(pred).requires_grad_()
    total, comps = pose.ComprehensivePoseLoss()(p, _t(gt))
    total.backward()
    got5 = np.array([comps[k].item() for k in ("mse_loss", "l1_loss", "inter_joint_loss", "abs_root_loss", "total_loss")])
    assert np.allclose(got5, ref5, rtol=1e-3, atol=0), (got5, ref5)   # north_star: 1e-3 relative on fp32 losses
    assert np.allclose(got5, ref5, rtol=2e-6), "fp32 kernel should be far inside the budget"
    g = p.grad.cpu().numpy()
    assert np.isfinite(g).all()
    assert np.abs(g - refg).max() <= 1e-5 * np.abs(refg).max()
    assert total.dim() == 0 and comps["mse_loss"].dim() == 0


def test_loss_matches_reference_golden(pose, golden):
    g = golden("loss.npz")
    for name in ("b8", "b1", "b33_w"):
        w = g[name + "_weights"]
        crit = pose.ComprehensivePoseLoss(mse_weight=float(w[0]), l1_weight=float(w[1]),
                                          inter_joint_loss_weight=float(w[2]), abs_root_loss_weight=float(w[3]))
        p = _t(g[name + "_pred"]).requires_grad_()
        total, comps = crit(p, _t(g[name + "_gt"]))
        (total * 0.5).backward()  # upstream gradient scaling goes through autograd
        got5 = np.array([comps[k].item() for k in ("mse_loss", "l1_loss", "inter_joint_loss", "abs_root_loss", "total_loss")])
        assert np.allclose(got5, g[name + "_out5"], rtol=1e-5)
        ref = g[name + "_grad"] * 0.5
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


def test_loss_workspace_reuse_and_determinism(pose):
    rng = np.random.default_rng(0)
    pred, gt = _loss_case(rng, 4096)
    crit = pose.ComprehensivePoseLoss()
    outs = []
    for _ in range(3):
        p = _t(pred).requires_grad_()
        total, _ = crit(p, _t(gt))
        total.backward()
        outs.append((total.item(), p.grad.clone()))
    assert outs[0][0] == outs[1][0] == outs[2][0]
    assert torch.equal(outs[0][1], outs[2][1])


def test_loss_other_joint_counts_and_errors(pose, oracle):
    rng = np.random.default_rng(1)
    for J in (1, 2, 32):
        pred, gt = _loss_case(rng, 5, J)
        ref5, refg = oracle.pose_loss(pred, gt)
        out5, grad = pose.loss.pose_loss_fwd_bwd(_t(pred), _t(gt), (1.0, 1.0, 100.0, 1.0))
        got = out5.cpu().numpy()
        if J == 1:  # mean over zero pairs is NaN in torch and in both implementations
            assert np.isnan(got[2]) and np.isnan(ref5[2])
            assert np.allclose(got[[0, 1, 3]], ref5[[0, 1, 3]], rtol=2e-6)
        else:
            assert np.allclose(got, ref5, rtol=2e-6)
            assert np.abs(grad.cpu().numpy() - refg).max() <= 1e-5 * np.abs(refg).max()
    with pytest.raises(RuntimeError):
        pose.loss.pose_loss_fwd_bwd(torch.zeros(2, 40, 3, device=DEV), torch.zeros(2, 40, 3, device=DEV), (1, 1, 1, 1))
    with pytest.raises(ValueError):
        pose.loss.pose_loss_fwd_bwd(torch.zeros(2, 17, 3, device=DEV), torch.zeros(3, 17, 3, device=DEV), (1, 1, 1, 1))


# --------------------------------------------------------------------------------------------- heatmap
def _kp_case(rng, B, J=17):
    kp = rng.uniform(0.02, 0.98, (B, J, 2)).astype(np.float32)
    m = rng.random((B, J)) < 0.05
    kp[m] = -1.0                 # invalid rows (<= 0) exercise the mask
    kp[0, 0] = [0.5, 0.5]
    if B > 1:
        kp[1, 1] = [1.0, 1.0]
        kp[1, 2] = [0.0, 0.4]   # exactly 0 is invalid (strict >)
    return kp


@pytest.mark.parametrize("hs,sigma,B", [(64, 2.0, 8), (256, 10.0, 4), (32, 1.5, 3), (50, 3.0, 2), (500, 10.0, 1)])
def test_heatmap_matches_oracle(pose, oracle, hs, sigma, B):
    kp = _kp_case(np.random.default_rng(hs), B)
    ref = oracle.heatmap(kp, hs, sigma)
    got = pose.GaussianHeatmapGenerator(17, hs, sigma).to(DEV)(_t(kp))
    assert got.shape == (B, 17, hs, hs) and got.dtype == torch.float32
    g = got.cpu().numpy()
    # arg-max location per plane: bit-exact (north_star)
    peak = oracle.heatmap_peak(kp, hs)
    am = g.reshape(B, 17, -1).argmax(-1)
    valid = peak >= 0
    assert np.array_equal(am[valid], peak[valid])
    assert (g[~valid] == 0).all()
    # values: CUDA expf vs glibc expf, <= 2 ulp on normal numbers
    assert np.allclose(g, ref, rtol=2.5e-7, atol=1e-37)


def test_heatmap_matches_reference_golden(pose, golden):
    g = golden("heatmap.npz")
    for name in ("s32", "vit", "cnn", "odd"):
        kp, hs, sigma = g[name + "_kp"], int(g[name + "_hs"]), float(g[name + "_sigma"])
        hm = pose.GaussianHeatmapGenerator(17, hs, sigma).to(DEV)(_t(kp)).cpu().numpy()
        B = kp.shape[0]
        assert np.array_equal(hm.reshape(B, 17, -1).argmax(-1), g[name + "_argmax"])
        ref, got = (g[name + "_full"], hm) if name + "_full" in g else (g[name + "_rows"], hm[:, :, :: max(1, hs // 8), :])
        assert np.allclose(got, ref, rtol=4e-7, atol=1e-37)
        assert np.allclose(hm.astype(np.float64).sum((2, 3)), g[name + "_sum"], rtol=1e-6)


def test_heatmap_nan_keypoint_and_layouts(pose, oracle):
    kp = _kp_case(np.random.default_rng(9), 3)
    hs, sigma = 64, 2.0
    ref = oracle.heatmap(kp, hs, sigma)
    common = pose.models.common
    # bf16 planes
    b16 = common.render_heatmaps(_t(kp), hs, sigma, dtype=torch.bfloat16).float().cpu().numpy()
    assert np.allclose(b16, ref, rtol=2 ** -8, atol=1e-30)
    # channels-last into the 21-channel conv operand (3 RGB + 1 depth + 17 heat-maps), fp32 and bf16
    for dt, tol in ((torch.float32, 2.5e-7), (torch.bfloat16, 2 ** -8)):
        x = torch.full((3, hs, hs, 24), -7.0, dtype=dt, device=DEV)
        common.render_heatmaps(_t(kp), hs, sigma, out=x, dtype=dt, channels_last=True, c_offset=4)
        xc = x.float().cpu().numpy()
        assert np.allclose(xc[..., 4:21].transpose(0, 3, 1, 2), ref, rtol=tol, atol=1e-30)
        assert (xc[..., :4] == -7.0).all() and (xc[..., 21:] == -7.0).all()
    # a NaN key-point gives a NaN plane (NaN * 0), like the reference
    kp2 = kp.copy()
    kp2[2, 5, 0] = np.nan
    hm = pose.GaussianHeatmapGenerator(17, hs, sigma).to(DEV)(_t(kp2)).cpu().numpy()
    assert np.isnan(hm[2, 5]).all() and np.isfinite(np.delete(hm.reshape(51, -1), 2 * 17 + 5, 0)).all()


def test_heatmap_full_size_properties(pose, oracle):
    """BASELINE config 2 size (B=256, hs=256): properties that do not need the oracle at full size."""
    B, hs, sigma = 256, 256, 10.0
    kp = _kp_case(np.random.default_rng(2), B)
    hm = pose.GaussianHeatmapGenerator(17, hs, sigma).to(DEV)(_t(kp))
    flat = hm.view(B, 17, -1)
    am = flat.argmax(-1).cpu().numpy()
    peak = oracle.heatmap_peak(kp, hs)
    valid = peak >= 0
    assert np.array_equal(am[valid], peak[valid])
    mx = flat.max(-1).values.cpu().numpy()
    assert (mx[~valid] == 0).all() and (mx[valid] <= 1.0).all() and (mx[valid] > 0.99).all()
    assert torch.isfinite(hm).all() and (hm >= 0).all()
    # spot-check 3 samples against the oracle
    idx = [0, 101, 255]
    assert np.allclose(hm[idx].cpu().numpy(), oracle.heatmap(kp[idx], hs, sigma), rtol=2.5e-7, atol=1e-37)


# --------------------------------------------------------------------------------------------- augment
def _aug_inputs(rng, B, H, W, root_relative=False):
    img = rng.random((B, 3, H, W), dtype=np.float32)
    dep = rng.random((B, 1, H, W), dtype=np.float32)
    img[:, :, : H // 4, : W // 3] = 0.5          # flat patch: integer-valued bilinear results
    img[:, 1, H // 2:, :] = np.float32(1.0)       # saturated plane
    kp = rng.uniform(0.05, 0.95, (B, 17, 2)).astype(np.float32)
    joints = rng.normal(0, 300, (B, 17, 3)).astype(np.float32)
    if not root_relative:
        joints[:, :, 2] += 4000
    cam = np.tile(np.array([1145.0 * W / 1000, 1144.0 * H / 1000, W / 2.0, H / 2.0]), (B, 1))
    return img, dep, kp, joints, cam


def _check_aug(pose, oracle, aug, img, dep, kp, joints, cam, params, pad_to=None):
    B = img.shape[0]
    out = aug.augment_batch(_t(img), _t(dep), _t(kp), _t(joints), _t(cam), params=params, pad_to=pad_to)
    assert aug.kernel_error_flag() == 0
    o_img, o_dep = out["image"].cpu().numpy(), out["depth"].cpu().numpy()
    sizes = out["sizes"].cpu().numpy()
    for i in range(B):
        ref = oracle.augment_sample(img[i], dep[i], kp[i], joints[i], cam[i], params[i], aug.flags)
        h, w = ref["image"].shape[1:]
        assert tuple(sizes[i]) == (h, w), (i, sizes[i], h, w)
        # uint8 pixels: north_star allows 1 LSB; this implementation is bit-exact
        assert np.array_equal(o_img[i, :, :h, :w], ref["image"]), f"sample {i}: RGB differs, params {params[i]}"
        assert np.array_equal(o_dep[i, :, :h, :w], ref["depth"]), f"sample {i}: depth differs"
        # collator-style zero padding
        assert (o_img[i, :, h:, :] == 0).all() and (o_img[i, :, :, w:] == 0).all()
        assert (o_dep[i, :, h:, :] == 0).all() and (o_dep[i, :, :, w:] == 0).all()
        # key-point / joint arithmetic: bit-exact fp32
        assert np.array_equal(out["keypoints_2d"][i].cpu().numpy().view(np.uint32), ref["keypoints_2d"].view(np.uint32)), i
        assert np.array_equal(out["joints_3d"][i].cpu().numpy().view(np.uint32), ref["joints_3d"].view(np.uint32)), i
        assert np.array_equal(out["camera"][i].cpu().numpy(), ref["cam"])
    return out


@pytest.mark.parametrize("H,W,B", [(64, 64, 16), (256, 256, 6), (48, 80, 8), (128, 128, 9)])
def test_augment_matches_oracle_all_stages(pose, oracle, H, W, B):
    rng = np.random.default_rng(H * 7 + W)
    img, dep, kp, joints, cam = _aug_inputs(rng, B, H, W, root_relative=(H == 128))
    aug = pose.PoseAugmentor()
    np.random.seed(H + W)
    params = aug.draw_params(B)
    params[0, 2] = 1.0     # unchanged size: Pillow's resize returns a copy
    params[1, 2] = 0.8
    params[2, 2] = 1.2
    _check_aug(pose, oracle, aug, img, dep, kp, joints, cam, params)


def test_augment_stage_subsets_and_extremes(pose, oracle):
    rng = np.random.default_rng(5)
    H = W = 96
    B = 6
    img, dep, kp, joints, cam = _aug_inputs(rng, B, H, W)
    ctors = [dict(enable_rotation=False, enable_scale=False), dict(enable_flip=False, enable_translate=False, enable_color=False),
             dict(flip_prob=1.0, enable_rotation=False, enable_scale=False, enable_translate=False, enable_color=False),
             dict(enable_flip=False, enable_rotation=False, enable_scale=True, enable_translate=False, enable_color=False),
             dict(brightness_range=(0.3, 1.9), contrast_range=(0.2, 2.5), rotation_range=(-180, 180), scale_range=(0.5, 1.7)),
             dict(enable_flip=False, enable_rotation=False, enable_scale=False, enable_translate=False, enable_color=False)]
    for kw in ctors:
        aug = pose.PoseAugmentor(**kw)
        np.random.seed(11)
        params = aug.draw_params(B)
        _check_aug(pose, oracle, aug, img, dep, kp, joints, cam, params)
    # exact multiples of 90 degrees take Pillow's transpose shortcuts; big translations leave only fill
    aug = pose.PoseAugmentor()
    np.random.seed(3)
    params = aug.draw_params(B)
    params[:, 1] = [0.0, 90.0, 180.0, -90.0, 360.0, 29.999]
    params[4, 3:5] = [0.99, -0.99]
    _check_aug(pose, oracle, aug, img, dep, kp, joints, cam, params, pad_to=(120, 120))


def test_augment_matches_reference_golden(pose, golden):
    g = golden("augment.npz")
    for i in range(int(g["n"])):
        k = f"c{i}_"
        aug = pose.PoseAugmentor(**json.loads(str(g[k + "ctor"])))
        img8, dep8 = g[k + "image_u8"], g[k + "depth_u8"]
        img = img8.astype(np.float32) / np.float32(255)
        dep = dep8.astype(np.float32) / np.float32(255)
        for as_u8 in (False, True):   # uint8 input reproduces the reference's p/255 -> *255 -> byte round trip exactly
            a = (_t(img8[None]), _t(dep8[None])) if as_u8 else (_t(img[None]), _t(dep[None]))
            out = aug.augment_batch(a[0], a[1], _t(g[k + "kp"][None]), _t(g[k + "joints"][None]), _t(g[k + "cam"][None]),
                                    params=g[k + "params"][None])
            h, w = out["sizes"][0].tolist()
            ref_i, ref_d = g[k + "out_image_u8"], g[k + "out_depth_u8"]
            assert (h, w) == ref_i.shape[1:], (i, h, w, ref_i.shape)
            got_i = out["image"][0, :, :h, :w].cpu().numpy()
            got_d = out["depth"][0, :, :h, :w].cpu().numpy()
            assert np.array_equal(got_i, ref_i.astype(np.float32) / np.float32(255)), f"golden case {i}"
            assert np.array_equal(got_d, ref_d.astype(np.float32) / np.float32(255)), f"golden case {i}"
            assert np.array_equal(out["keypoints_2d"][0].cpu().numpy().view(np.uint32), g[k + "out_kp"].view(np.uint32))
            assert np.array_equal(out["joints_3d"][0].cpu().numpy().view(np.uint32), g[k + "out_joints"].view(np.uint32))
            assert np.array_equal(out["camera"][0].cpu().numpy(), g[k + "out_cam"])


def test_augment_call_is_drop_in(pose, oracle):
    """__call__(sample) -> sample, CPU tensors in and out like the reference, global np.random consumed in order."""
    rng = np.random.default_rng(8)
    img, dep, kp, joints, cam = _aug_inputs(rng, 1, 64, 64)
    sample = dict(image=torch.from_numpy(img[0]), depth=torch.from_numpy(dep[0]), keypoints_2d=torch.from_numpy(kp[0]),
                  joints_3d=torch.from_numpy(joints[0]), camera_params=dict(R=None, t=None, f=list(cam[0, :2]), c=list(cam[0, 2:])),
                  extra="kept")
    aug = pose.PoseAugmentor()
    np.random.seed(21)
    params = aug.draw_params(1)
    np.random.seed(21)
    out = aug(sample)
    ref = oracle.augment_sample(img[0], dep[0], kp[0], joints[0], cam[0], params[0], aug.flags)
    assert out["extra"] == "kept" and out["image"].device.type == "cpu"
    assert np.array_equal(out["image"].numpy(), ref["image"]) and np.array_equal(out["depth"].numpy(), ref["depth"])
    assert np.array_equal(out["keypoints_2d"].numpy(), ref["keypoints_2d"])
    assert out["camera_params"]["f"] == list(ref["cam"][:2]) and sample["camera_params"]["f"] == list(cam[0, :2])


def test_augment_full_size_properties(pose, oracle):
    """BASELINE config 2 size (B=256, 256x256): identity parameters round-trip; flip twice is the identity;
    random samples spot-checked against the oracle."""
    B, H, W = 256, 256, 256
    rng = np.random.default_rng(4)
    img, dep, kp, joints, cam = _aug_inputs(rng, B, H, W)
    aug = pose.PoseAugmentor()
    ident = np.zeros((B, 8))
    ident[:, 2] = 1.0
    ident[:, 5:7] = 1.0
    out = aug.augment_batch(_t(img), _t(dep), _t(kp), _t(joints), _t(cam), params=ident)
    q = lambda a: (a * np.float32(255)).astype(np.uint8).astype(np.float32) / np.float32(255)
    assert out["image"].shape == (B, 3, 256, 256)
    assert np.array_equal(out["image"].cpu().numpy(), q(img)) and np.array_equal(out["depth"].cpu().numpy(), q(dep))
    flip = ident.copy()
    flip[:, 0] = 1.0
    o1 = aug.augment_batch(_t(img), _t(dep), _t(kp), _t(joints), _t(cam), params=flip)
    assert np.array_equal(o1["image"].cpu().numpy(), q(img)[..., ::-1])
    np.random.seed(0)
    params = aug.draw_params(B)
    out = aug.augment_batch(_t(img), _t(dep), _t(kp), _t(joints), _t(cam), params=params)
    assert aug.kernel_error_flag() == 0
    sizes = out["sizes"].cpu().numpy()
    assert np.array_equal(sizes[:, 0], (256 * params[:, 2]).astype(int))
    for i in (0, 17, 128, 255):
        ref = oracle.augment_sample(img[i], dep[i], kp[i], joints[i], cam[i], params[i], aug.flags)
        h, w = ref["image"].shape[1:]
        assert np.array_equal(out["image"][i, :, :h, :w].cpu().numpy(), ref["image"])
        assert np.array_equal(out["depth"][i, :, :h, :w].cpu().numpy(), ref["depth"])
        assert np.array_equal(out["keypoints_2d"][i].cpu().numpy(), ref["keypoints_2d"])


# ------------------------------------------------------------------------------ tcgen05 GEMM / head
def _ref_linear(a_bf16, w_bf16, bias, act):
    """Plain PyTorch fp32 reference of the same op on the same bf16-rounded operands."""
    y = a_bf16.float() @ w_bf16.float().t()
    if bias is not None:
        y = y + bias
    if act == "silu":
        y = torch.nn.functional.silu(y)
    elif act == "gelu":
        y = torch.nn.functional.gelu(y)
    elif act == "relu":
        y = torch.relu(y)
    return y


@pytest.mark.parametrize("M,N,K,act", [(256, 1024, 1024, "silu"), (128, 128, 64, None), (1, 51, 512, None),
                                       (300, 200, 136, "gelu"), (257, 768, 3072, "relu"), (4096, 64, 64, None),
                                       (130, 16, 512, None), (1000, 3072, 768, "gelu"),
                                       # 128 x 256 tiles (N % 256 == 0 and >= 148 tiles), ragged last M tile
                                       (16448, 768, 768, "silu"), (6500, 2304, 320, None)])
def test_gemm_tcgen05_matches_fp32_reference(pose, M, N, K, act):
    g = torch.Generator(device="cpu").manual_seed(M * 31 + N)
    a = (torch.randn(M, K, generator=g) * 0.5).to(DEV).bfloat16()
    w = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).to(DEV).bfloat16()
    bias = torch.randn(N, generator=g).to(DEV)
    ops = pose.ops
    ref = _ref_linear(a, w, bias, act)
    got = ops.gemm_bf16(a, w, bias, act=act, out_dtype=torch.float32)
    # fp32 accumulation of exact bf16 products: only summation order differs
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (got - ref).abs().max().item()
    got16 = ops.gemm_bf16(a, w, None, act=None, out_dtype=torch.bfloat16)
    ref16 = _ref_linear(a, w, None, None)
    assert torch.allclose(got16.float(), ref16, rtol=2 ** -7, atol=1e-2)


def test_regression_head_matches_fp32_reference(pose):
    """PoseRegressionHead (common.py:55-89) CNN flavour 1024-1024-512-51 SiLU and ViT-style dims with GELU.
    Tolerance: bf16 operands / fp32 accumulate against the fp32 module on outputs of O(1)."""
    for in_f, hidden, act in ((1024, [1024, 512], "silu"), (768, [1024, 512, 256], "gelu")):
        torch.manual_seed(0)
        head = pose.PoseRegressionHead(in_f, 17, hidden_dims=hidden, dropout=0.2, activation=act).to(DEV).eval()
        x = torch.randn(256, in_f, device=DEV)
        with torch.no_grad():
            got = head(x)
            h = x
            lins = [m[0] if isinstance(m, torch.nn.Sequential) else m for m in head.decoder]
            for i, lin in enumerate(lins):
                h = torch.nn.functional.linear(h, lin.weight, lin.bias)
                if i < len(lins) - 1:
                    h = torch.nn.functional.silu(h) if act == "silu" else torch.nn.functional.gelu(h)
            ref = h.view(-1, 17, 3)
        assert got.shape == (256, 17, 3) and got.dtype == torch.float32
        err = (got - ref).abs().max().item()
        assert err < 3e-2 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("flavour", ["common", "transformers"])
def test_regression_head_training_mode_autograd_and_dropout(pose, flavour):
    """Standalone PoseRegressionHead in training mode (common.py:69-89 / transformers.py:7-31): dropout + autograd through
    the module itself.  The counter-based mask is exposed through pose_dropout_bf16 (same seed, tensor of ones), so the
    fused forward / backward can be compared with plain fp32 PyTorch using that very mask.  Tolerance: bf16 operands."""
    import importlib
    ops = importlib.import_module("3dhumanposeestimation_b200.ops")
    lib = pose._lib.lib()
    torch.manual_seed(3)
    if flavour == "common":
        head = pose.PoseRegressionHead(1024, 17, hidden_dims=[1024, 512], dropout=0.2, activation="silu").to(DEV).train()
        lins = [m[0] if isinstance(m, torch.nn.Sequential) else m for m in head.decoder]
        p, act = 0.2, torch.nn.functional.silu
    else:
        tf = importlib.import_module("3dhumanposeestimation_b200.models.transformers")
        head = tf.PoseRegressionHead(768, 17, hidden_dims=[1024, 512, 256], dropout=0.25, activation="gelu").to(DEV).train()
        lins = head.linears()
        p, act = 0.25, torch.nn.functional.gelu
    B = 192
    x = torch.randn(B, lins[0].in_features, device=DEV, requires_grad=True)
    seed = ((torch.initial_seed() & 0xFFFFFFFF) << 20) + (ops._head_calls[0] + 1) * 16     # the seed the call will draw
    out = head(x)
    assert out.shape == (B, 17, 3) and out.dtype == torch.float32 and out.requires_grad
    w = torch.randn_like(out)
    (out * w).sum().backward()
    got = {"x": x.grad.clone()}
    for i, lin in enumerate(lins):
        got[f"w{i}"], got[f"b{i}"] = lin.weight.grad.clone(), lin.bias.grad.clone()
        lin.weight.grad = lin.bias.grad = None
    # plain PyTorch with the same masks
    xr = x.detach().clone().requires_grad_()
    h = xr
    for i, lin in enumerate(lins):
        h = torch.nn.functional.linear(h, lin.weight, lin.bias)
        if i < len(lins) - 1:
            ones = torch.ones(B, lin.out_features, device=DEV, dtype=torch.bfloat16)
            mask = torch.empty_like(ones)
            rc = lib.pose_dropout_bf16(ones.data_ptr(), ones.numel(), p, seed + i, mask.data_ptr(), pose._lib.stream_ptr())
            assert rc == 0
            keep = (mask != 0).float()
            assert abs(keep.mean().item() - (1 - p)) < 0.02
            h = act(h) * keep / (1 - p)
    ref = h.view(B, 17, 3)
    (ref * w).sum().backward()
    assert (out - ref).abs().max().item() < 3e-2 * max(1.0, ref.abs().max().item())

    def rel(a, b):
        return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
    errs = {"x": rel(got["x"], xr.grad)}
    for i, lin in enumerate(lins):
        errs[f"w{i}"] = rel(got[f"w{i}"], lin.weight.grad)
        errs[f"b{i}"] = rel(got[f"b{i}"], lin.bias.grad)
    assert max(errs.values()) < 2e-2, errs
    # eval mode: no dropout, deterministic, no autograd graph needed
    head.eval()
    with torch.no_grad():
        a, b = head(x.detach()), head(x.detach())
    assert torch.equal(a, b)


def test_eval_metrics_match_the_reference(pose, oracle, golden):
    """MPJPE / PA-MPJPE on the device (one thread per sample, closed-form 3x3 SVD) vs the live reference's values and the
    C oracle, incl. the reflection, similarity-transform, identical-pose and collapsed-prediction cases; 1e-3 relative."""
    d = golden("metrics.npz")
    pred, gt = torch.from_numpy(d["pred"]).cuda(), torch.from_numpy(d["gt"]).cuda()
    means, per = pose.utils.eval_metrics(pred, gt)
    assert abs(means[0].item() - float(d["mpjpe"])) < 1e-3 * float(d["mpjpe"])
    assert abs(means[1].item() - float(d["pa_mpjpe"])) < 1e-3 * float(d["pa_mpjpe"])
    assert np.allclose(per[1].cpu().numpy(), d["pa_per_sample"], rtol=1e-4, atol=2e-3)
    assert abs(pose.utils.compute_pa_mpjpe(pred, gt).item() - oracle.pa_mpjpe(d["pred"], d["gt"])) < 1e-2
    # a large seeded batch against the oracle
    rng = np.random.default_rng(3)
    g2 = rng.normal(0, 300, (4096, 17, 3)).astype(np.float32)
    p2 = (g2 + rng.normal(0, 80, g2.shape)).astype(np.float32)
    m2, per2 = pose.utils.eval_metrics(torch.from_numpy(p2).cuda(), torch.from_numpy(g2).cuda())
    mo, pero = oracle.pa_mpjpe(p2, g2, per_sample=True)
    assert np.allclose(per2[1].cpu().numpy(), pero, rtol=1e-4, atol=1e-3)
    assert abs(m2[0].item() - oracle.mpjpe(p2, g2)) < 1e-3 and abs(m2[1].item() - mo) < 1e-3


@pytest.mark.parametrize("h,w,H,W", [(37, 53, 256, 256), (480, 640, 256, 256), (256, 256, 256, 256), (1, 7, 5, 9),
                                     (300, 200, 512, 512)])
def test_infer_prep_matches_oracle_and_torch(pose, h, w, H, W):
    """SURVEY 8f rank 4 (infer.py:217-221, :362-367): depth resize bit-exact against the numpy oracle (same op order, no
    contraction), within 2 ulp-ish of torch's own F.interpolate; key-point normalisation bit-exact."""
    import importlib
    import torch.nn.functional as F
    from oracle import infer_ref
    infer = importlib.import_module("3dhumanposeestimation_b200.infer")
    rng = np.random.default_rng(h * 7 + W)
    d = rng.random((3, h, w), dtype=np.float32) * 10.0
    k = rng.random((3, 17, 3), dtype=np.float32) * np.float32(max(h, w))
    out, kp2, kp3 = infer.prepare_model_inputs(torch.from_numpy(d)[:, None].to(DEV), torch.from_numpy(k).to(DEV), (w, h), (H, W))
    assert out.shape == (3, 1, H, W) and kp2.shape == (3, 17, 2) and kp3.shape == (3, 17, 3)
    assert np.array_equal(out[:, 0].cpu().numpy(), infer_ref.depth_resize(d, H, W))
    want = F.interpolate(torch.from_numpy(d)[:, None], size=(H, W), mode="bilinear", align_corners=False)
    assert torch.allclose(out.cpu(), want, rtol=2e-6, atol=2e-6)
    r2, r3 = infer_ref.normalise_keypoints(k, w, h)
    assert np.array_equal(kp2.cpu().numpy(), r2) and np.array_equal(kp3.cpu().numpy(), r3)
    with pytest.raises(Exception):
        infer.prepare_model_inputs(torch.from_numpy(d)[:, None], None, (w, h), (H, W))      # CPU tensor: no fallback


def test_run_inference_mirrors_the_reference_call(pose):
    """infer.py:383-393: eval-mode forward under no_grad, first sample as numpy."""
    import importlib
    infer = importlib.import_module("3dhumanposeestimation_b200.infer")
    cfg = pose.ModelConfig("cnn", image_size=(64, 64), heatmap_size=64, initial_channels=32, stage_channels=[64, 128, 256],
                           global_pool_size=2, global_feature_dim=256, regression_dims=[128, 64])
    m = pose.CNNPoseEstimation(cfg).to(DEV)
    g = torch.Generator().manual_seed(1)
    img = torch.rand(2, 3, 64, 64, generator=g).to(DEV)
    depth_raw = torch.rand(2, 1, 48, 80, generator=g).to(DEV)
    kpx = torch.rand(2, 17, 3, generator=g).to(DEV) * 80
    dep, kp, _ = infer.prepare_model_inputs(depth_raw, kpx, (80, 48), (64, 64))
    out = infer.run_inference(m, img, dep, kp)
    assert out.shape == (17, 3) and np.isfinite(out).all() and not m.training
    with torch.no_grad():
        assert np.array_equal(out, m(img, dep, kp)[0].float().cpu().numpy())


def _collate_batch(sizes, seed):
    """The seeded sample list oracle/gen_golden.py::collate_batch fed to the live reference collator."""
    g = torch.Generator().manual_seed(seed)
    batch = []
    for i, (h, w) in enumerate(sizes):
        batch.append({"image": torch.rand(3, h, w, generator=g), "depth": torch.rand(1, h, w, generator=g),
                      "keypoints_2d": torch.rand(17, 2, generator=g), "joints_3d": torch.randn(17, 3, generator=g),
                      "camera_params": {"f": [1.0 + i, 1.0]}, "image_path": f"p{i}", "action": "a", "subaction": i % 2 + 1,
                      "image_size": torch.tensor([h, w]), "frame_idx": i})
    return batch


def test_collator_matches_the_live_reference_golden(pose, golden):
    """Human36MCollator on the device against the LIVE reference collator's outputs
    (src/dataset/collator.py:10-61, frozen by oracle/gen_golden.py::gen_collate): ragged sizes, a single sample, equal
    sizes -- every field of the result dictionary, bit-exact."""
    import importlib
    col = importlib.import_module("3dhumanposeestimation_b200.dataset.collator").Human36MCollator()
    gold = golden("collate.npz")
    for c in range(int(gold["n_cases"])):
        sizes = [tuple(int(v) for v in row) for row in gold[f"c{c}_sizes"]]
        batch = [{k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in smp.items()}
                 for smp in _collate_batch(sizes, 900 + c)]
        out = col(batch)
        for key in ("image", "depth", "keypoints_2d", "joints_3d", "image_size"):
            want = gold[f"c{c}_{key}"]
            got = out[key].cpu().numpy()
            assert got.shape == want.shape and got.dtype == want.dtype, (c, key, got.shape, got.dtype)
            assert np.array_equal(got, want), (c, key)
        assert [tuple(p) for p in out["padding"]] == [tuple(int(v) for v in row) for row in gold[f"c{c}_padding"]]
        lists = json.loads(str(gold[f"c{c}_lists"]))
        for key, want in lists.items():
            assert out[key] == want, (c, key)
        # the dataset's depth rescale (chunked_dataset.py:159-164) fused into the same launch: same torch ops as the reference
        rng = [(0.5 * i, 0.5 * i + 3.0) for i in range(len(sizes))]
        out2 = col(batch, depth_range=rng)
        mh, mw = out["padding"][0]
        want = torch.stack([torch.nn.functional.pad(smp["depth"] * (hi - lo) + lo, (0, mw - smp["depth"].shape[2], 0, mh - smp["depth"].shape[1]))
                            for smp, (lo, hi) in zip(batch, rng)])
        assert torch.equal(out2["depth"], want)
    with pytest.raises(Exception):
        col([{**batch[0], "image": batch[0]["image"].cpu()}])          # CPU tensors: no fallback


def test_resize_of_decoded_frames_matches_the_reference_bit_for_bit(pose, oracle, golden):
    """transforms.Resize(image_size) + depth rescale on the device (pose_resize_bilinear_aa) against the
    live torchvision outputs frozen in tests/golden/resize.npz and against the C oracle at the dataset's real frame size
    (1000 x 1002 -> 256 x 256), uint8 and float32 inputs, batched.  Bit-exact."""
    import importlib
    tr = importlib.import_module("3dhumanposeestimation_b200.dataset.transforms")
    gold = golden("resize.npz")
    for i in range(int(gold["n_cases"])):
        c, h, w, oh, ow = (int(v) for v in gold["cases"][i])
        u8 = np.random.default_rng(700 + i).integers(0, 256, (c, h, w), dtype=np.uint8)
        got8 = tr.Resize((oh, ow))(torch.from_numpy(u8).to(DEV))
        # float32 frames as the reference builds them on the CPU (`.float() / 255.0`: an IEEE division there; torch's CUDA
        # division by a scalar multiplies by the reciprocal instead, so the conversion is done before the copy)
        gotf = tr.resize_frames((torch.from_numpy(u8).float() / 255.0).to(DEV), (oh, ow))
        assert got8.dtype == torch.float32 and tuple(got8.shape) == (c, oh, ow)
        assert np.array_equal(got8.cpu().numpy(), gold[f"r{i}"]), (i, np.abs(got8.cpu().numpy() - gold[f"r{i}"]).max())
        assert np.array_equal(gotf.cpu().numpy(), gold[f"r{i}"])
        if c == 1:
            lo, hi = (float(v) for v in gold[f"r{i}_range"])
            gd = tr.resize_frames(torch.from_numpy(u8).to(DEV)[None], (oh, ow), depth_range=[(lo, hi)])[0]
            assert np.array_equal(gd.cpu().numpy(), gold[f"r{i}_depth"])
    # the dataset's frame size, a batch of frames in one launch
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (3, 3, 1000, 1002), dtype=np.uint8)
    out = tr.resize_frames(torch.from_numpy(frames).to(DEV), (256, 256)).cpu().numpy()
    for b in range(3):
        ref = oracle.tensor_resize_aa(frames[b].astype(np.float32) / np.float32(255.0), 256, 256)
        assert np.array_equal(out[b], ref), (b, np.abs(out[b] - ref).max())
    big = tr.resize_frames(torch.from_numpy(frames[:1]).to(DEV), (500, 500)).cpu().numpy()[0]
    assert np.array_equal(big, oracle.tensor_resize_aa(frames[0].astype(np.float32) / np.float32(255.0), 500, 500))
    with pytest.raises(Exception):
        tr.Resize(64)(torch.rand(3, 80, 80))            # CPU tensor: no fallback
