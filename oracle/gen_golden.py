"""Generates tests/golden/*.npz from the LIVE reference (run in this container only).

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/

The reference is pure Python and cannot travel to the GPU box, so its outputs on small seeded inputs
are frozen here.  Library versions are recorded in every file (the reference pins torchvision 0.19.0 /
Pillow 11.1.0 / numpy 1.26.4; this container has newer ones -- see `versions`).
The script runs from a scratch directory because importing the reference's `config` creates
./logs and ./dataset_cache (src/config.py:41-45).
"""
import json
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "tests", "golden")
REF = os.environ.get("POSE_REFERENCE", "/root/reference")


def versions():
    import PIL
    import torch
    import torchvision
    return json.dumps({"torch": torch.__version__, "torchvision": torchvision.__version__, "pillow": PIL.__version__,
                       "numpy": np.__version__})


def make_sample(rng, H, W, root_relative):
    import torch
    img8 = rng.integers(0, 256, (3, H, W), dtype=np.uint8)
    dep8 = rng.integers(0, 256, (1, H, W), dtype=np.uint8)
    # smooth half of the image so the bilinear / antialias paths see flat and graded regions too
    yy, xx = np.mgrid[0:H, 0:W]
    img8[:, :, : W // 2] = ((yy[:, : W // 2] * 3 + xx[:, : W // 2] * 2) % 256).astype(np.uint8)[None]
    img8[1, : H // 3, :] = 255
    kp = rng.uniform(0.05, 0.95, (17, 2)).astype(np.float32)
    joints = rng.normal(0, 300, (17, 3)).astype(np.float32)
    if not root_relative:
        joints[:, 2] += 4000
    cam = [1145.0 * W / 1000, 1144.0 * H / 1000, W / 2.0, H / 2.0]
    sample = dict(image=torch.from_numpy(img8.astype(np.float32) / np.float32(255)),
                  depth=torch.from_numpy(dep8.astype(np.float32) / np.float32(255)),
                  keypoints_2d=torch.from_numpy(kp), joints_3d=torch.from_numpy(joints),
                  camera_params=dict(R=None, t=None, f=cam[:2], c=cam[2:]))
    return sample, img8, dep8, kp, joints, np.array(cam, np.float64)


def gen_augment():
    from dataset.augmentation import PoseAugmentor
    cases = [
        (0, 64, 64, {}), (1, 64, 64, {}), (2, 64, 64, {}), (3, 64, 64, {}),
        (4, 48, 80, {}), (5, 80, 48, {}),
        (6, 64, 64, dict(enable_rotation=False, enable_scale=False)),
        (7, 64, 64, dict(enable_flip=False, enable_translate=False, enable_color=False)),
        (8, 64, 64, dict(flip_prob=1.0, enable_rotation=False, enable_scale=False, enable_translate=False,
                         enable_color=False)),
        (9, 64, 64, dict(brightness_range=(0.3, 1.9), contrast_range=(0.2, 2.5), rotation_range=(-180, 180),
                         scale_range=(0.5, 1.7))),
        (10, 64, 64, dict(enable_flip=False, enable_rotation=False, enable_scale=True, enable_translate=False,
                          enable_color=False)),
        (11, 256, 256, {}),
    ]
    out = {"versions": versions(), "n": len(cases)}
    for idx, (seed, H, W, kw) in enumerate(cases):
        rng = np.random.default_rng(1000 + seed)
        sample, img8, dep8, kp, joints, cam = make_sample(rng, H, W, root_relative=(seed % 3 == 2))
        aug = PoseAugmentor(**kw)
        # replay the global-RNG draws in the reference's order to record the explicit parameters
        np.random.seed(seed)
        p = np.zeros(8)
        if aug.enable_flip:
            p[0] = float(np.random.random() < aug.flip_prob)
        if aug.enable_rotation:
            p[1] = np.random.uniform(*aug.rotation_range)
        if aug.enable_scale:
            p[2] = np.random.uniform(*aug.scale_range)
        if aug.enable_translate:
            p[3] = np.random.uniform(*aug.translate_range)
            p[4] = np.random.uniform(*aug.translate_range)
        if aug.enable_color:
            p[5] = np.random.uniform(*aug.brightness_range)
            p[6] = np.random.uniform(*aug.contrast_range)
        np.random.seed(seed)
        ref = aug(sample)
        flags = (1 * aug.enable_flip) | (2 * aug.enable_rotation) | (4 * aug.enable_scale) | \
                (8 * aug.enable_translate) | (16 * aug.enable_color)
        oi = ref["image"].numpy()
        od = ref["depth"].numpy()
        oi8 = np.round(oi * 255).astype(np.uint8)
        od8 = np.round(od * 255).astype(np.uint8)
        assert np.array_equal(oi8.astype(np.float32) / np.float32(255), oi)
        assert np.array_equal(od8.astype(np.float32) / np.float32(255), od)
        cp = ref["camera_params"]
        k = f"c{idx}_"
        out[k + "seed"] = seed
        out[k + "ctor"] = json.dumps(kw)
        out[k + "flags"] = flags
        out[k + "params"] = p
        out[k + "image_u8"] = img8
        out[k + "depth_u8"] = dep8
        out[k + "kp"] = kp
        out[k + "joints"] = joints
        out[k + "cam"] = cam
        out[k + "out_image_u8"] = oi8
        out[k + "out_depth_u8"] = od8
        out[k + "out_kp"] = ref["keypoints_2d"].numpy()
        out[k + "out_joints"] = ref["joints_3d"].numpy()
        out[k + "out_cam"] = np.array(list(cp["f"]) + list(cp["c"]), np.float64)
    np.savez_compressed(os.path.join(OUT, "augment.npz"), **out)


def gen_heatmap():
    import torch
    from models.common import GaussianHeatmapGenerator
    rng = np.random.default_rng(7)
    out = {"versions": versions()}
    for name, hs, sigma, B in [("s32", 32, 1.5, 2), ("vit", 64, 2.0, 3), ("cnn", 256, 10.0, 2), ("odd", 50, 3.0, 2)]:
        kp = rng.uniform(0.02, 0.98, (B, 17, 2)).astype(np.float32)
        kp[0, 3] = [-1, -1]
        kp[0, 5, 0] = 0.0
        kp[1, 2] = [0.5, 0.5]
        kp[1, 1] = [1.0, 1.0]
        kp[1, 7] = [1e-6, 0.3]
        hm = GaussianHeatmapGenerator(17, hs, sigma)(torch.from_numpy(kp)).numpy()
        out[name + "_kp"] = kp
        out[name + "_hs"] = hs
        out[name + "_sigma"] = sigma
        out[name + "_argmax"] = hm.reshape(B, 17, -1).argmax(-1).astype(np.int32)
        out[name + "_sum"] = hm.astype(np.float64).sum((2, 3))
        out[name + "_max"] = hm.max((2, 3))
        if hs <= 32:
            out[name + "_full"] = hm
        else:
            out[name + "_rows"] = hm[:, :, :: max(1, hs // 8), :]  # every (hs/8)-th row
    np.savez_compressed(os.path.join(OUT, "heatmap.npz"), **out)


def gen_loss():
    import torch
    from loss import ComprehensivePoseLoss
    rng = np.random.default_rng(11)
    out = {"versions": versions()}
    keys = ["mse_loss", "l1_loss", "inter_joint_loss", "abs_root_loss", "total_loss"]
    for name, B, w in [("b8", 8, None), ("b1", 1, None), ("b33_w", 33, dict(l1_weight=0.5, mse_weight=2.0,
                                                                          inter_joint_loss_weight=10.0,
                                                                          abs_root_loss_weight=3.0))]:
        gt = rng.normal(0, 300, (B, 17, 3)).astype(np.float32)
        pred = (gt + rng.normal(0, 50, (B, 17, 3))).astype(np.float32)
        pred[0, 3] = pred[0, 4]      # coincident predicted joints: zero gradient through the norm
        pred[0, 7] = gt[0, 7]        # zero difference: sign(0) = 0
        p = torch.from_numpy(pred).requires_grad_()
        crit = ComprehensivePoseLoss(**(w or {}))
        total, comps = crit(p, torch.from_numpy(gt))
        total.backward()
        out[name + "_pred"] = pred
        out[name + "_gt"] = gt
        out[name + "_weights"] = np.array([crit.mse_weight, crit.l1_weight, crit.inter_joint_loss_weight,
                                           crit.abs_root_loss_weight], np.float32)
        out[name + "_out5"] = np.array([comps[k].item() for k in keys], np.float32)
        out[name + "_grad"] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)


def gen_config():
    from model_config import ModelConfig
    out = {"cnn": ModelConfig("cnn").to_dict(), "transformer": ModelConfig("transformer").to_dict(),
           "cnn_256": ModelConfig("cnn", image_size=(256, 256), heatmap_size=256).to_dict()}
    with open(os.path.join(OUT, "model_config.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


SMALL_CNN = dict(image_size=(64, 64), heatmap_size=64, heatmap_sigma=3.0, initial_channels=32,
                 stage_channels=[64, 128, 256], global_pool_size=2, global_feature_dim=256, regression_dims=[128, 64])


def gen_cnn():
    """State-dict layout of the full-size reference CNN + outputs of a small configuration (eval mode) whose
    parameters come from oracle.torch_models.fill_state_dict (deterministic, regenerated by the tests)."""
    import torch
    from model_config import ModelConfig
    from models.cnn import CNNPoseEstimation
    sys.path.insert(0, REPO)
    from oracle import torch_models as tm
    full = CNNPoseEstimation(ModelConfig("cnn", image_size=(256, 256), heatmap_size=256))
    layout = {k: list(v.shape) for k, v in full.state_dict().items()}
    with open(os.path.join(OUT, "cnn_state_dict_layout.json"), "w") as f:
        json.dump(layout, f, indent=0, sort_keys=True)
    cfg = ModelConfig("cnn", **SMALL_CNN)
    m = CNNPoseEstimation(cfg)
    sd = tm.fill_state_dict(m.state_dict(), seed=3)
    m.load_state_dict(sd)
    m.eval()
    g = torch.Generator().manual_seed(5)
    img = torch.rand(4, 3, 64, 64, generator=g)
    dep = torch.rand(4, 1, 64, 64, generator=g)
    kp = torch.rand(4, 17, 2, generator=g) * 0.9 + 0.05
    kp[0, 3] = -1.0
    with torch.no_grad():
        out = m(img, dep, kp)
        out_oracle = tm.cnn_forward(sd, cfg, img, dep, kp)
    assert (out - out_oracle).abs().max() < 1e-3
    np.savez_compressed(os.path.join(OUT, "cnn_small.npz"), versions=versions(), config=json.dumps(SMALL_CNN),
                        fill_seed=3, image=img.numpy(), depth=dep.numpy(), kp=kp.numpy(), out=out.numpy(),
                        n_params=sum(p.numel() for p in m.parameters()))


def _install_timm_stub():
    """timm is a third-party dependency of src/models/transformers.py:5 that is not under /root/reference and not
    installed here.  This stub restates timm 1.0.15's `VisionTransformer` (vit_base_patch16: conv
    patch embed, cls token + pos_embed, 12 pre-LN blocks with LayerNorm eps 1e-6, fused qkv, 12 heads, exact GELU MLP,
    final LayerNorm; `forward_features` returns all 1 + N tokens) with the attribute surface the reference touches,
    so that the reference's OWN TransformerPoseEstimation (patch-embed surgery, fusion blocks, final encoder, head)
    runs live on top of it."""
    import types
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    class Attention(nn.Module):
        def __init__(self, dim, heads):
            super().__init__()
            self.num_heads = heads
            self.qkv = nn.Linear(dim, dim * 3)
            self.proj = nn.Linear(dim, dim)

        def forward(self, x):
            B, N, C = x.shape
            qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
            x = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
            return self.proj(x.transpose(1, 2).reshape(B, N, C))

    class Mlp(nn.Module):
        def __init__(self, dim):
            super().__init__()
            self.fc1 = nn.Linear(dim, dim * 4)
            self.act = nn.GELU()
            self.fc2 = nn.Linear(dim * 4, dim)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    class Block(nn.Module):
        def __init__(self, dim, heads):
            super().__init__()
            self.norm1 = nn.LayerNorm(dim, eps=1e-6)
            self.attn = Attention(dim, heads)
            self.norm2 = nn.LayerNorm(dim, eps=1e-6)
            self.mlp = Mlp(dim)

        def forward(self, x):
            x = x + self.attn(self.norm1(x))
            return x + self.mlp(self.norm2(x))

    class PatchEmbed(nn.Module):
        def __init__(self, img_size, patch, dim):
            super().__init__()
            self.num_patches = (img_size[0] // patch) * (img_size[1] // patch)
            self.proj = nn.Conv2d(3, dim, patch, patch)

        def forward(self, x):
            return self.proj(x).flatten(2).transpose(1, 2)

    class VisionTransformer(nn.Module):
        default_cfg = {}          # no "embed_dim" key: the reference falls into its except branch (:164-170)

        def __init__(self, img_size, dim=768, depth=12, heads=12, patch=16):
            super().__init__()
            self.num_prefix_tokens = 1
            self.patch_embed = PatchEmbed(img_size, patch, dim)
            self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
            self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, dim))
            self.blocks = nn.Sequential(*[Block(dim, heads) for _ in range(depth)])
            self.norm = nn.LayerNorm(dim, eps=1e-6)

        def forward_features(self, x):
            x = self.patch_embed(x)
            x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], 1) + self.pos_embed
            return self.norm(self.blocks(x))

    def create_model(name, pretrained=False, num_classes=0, img_size=None, **kw):
        assert name.startswith("vit_base_patch16") and not pretrained
        return VisionTransformer(tuple(img_size) if img_size is not None else (384, 384))

    mod = types.ModuleType("timm")
    mod.create_model = create_model
    sys.modules["timm"] = mod


VIT_CFG = dict(image_size=(256, 256), vit_pretrained=False)
VIT_TRAIN_CFG = dict(image_size=(256, 256), vit_pretrained=False, transformer_dropout_rate=0.0,
                     transformer_attention_dropout_rate=0.0, regression_dropout=0.0)


def gen_vit():
    """Reference TransformerPoseEstimation (live, over the timm stub) at 256x256: eval-mode outputs, and one training
    step's loss + per-parameter gradient norms (dropout rates 0 so that the step is deterministic).  Parameters come from
    oracle.torch_models.fill_vit_state_dict (regenerated by the tests from the same seed)."""
    import contextlib
    import io
    import torch
    _install_timm_stub()
    from model_config import ModelConfig
    from models.transformers import TransformerPoseEstimation
    from loss import ComprehensivePoseLoss
    sys.path.insert(0, REPO)
    from oracle import torch_models as tm
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = TransformerPoseEstimation(ModelConfig("transformer", **VIT_CFG))
    layout = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, "vit_state_dict_layout.json"), "w") as f:
        json.dump(layout, f, indent=0, sort_keys=True)
    sd = tm.fill_vit_state_dict(m.state_dict(), seed=7)
    m.load_state_dict(sd)
    m.eval()
    g = torch.Generator().manual_seed(9)
    img = torch.rand(2, 3, 256, 256, generator=g)
    dep = torch.rand(2, 1, 256, 256, generator=g)
    kp = torch.rand(2, 17, 2, generator=g) * 0.9 + 0.05
    kp[1, 5] = -1.0
    with torch.no_grad():
        out = m(img, dep, kp)
        out_oracle = tm.vit_forward(sd, m.config, img, dep, kp)
    assert (out - out_oracle).abs().max() < 2e-3 * out.abs().max(), (out - out_oracle).abs().max()
    # one training step (dropout 0): loss and gradient norms
    with contextlib.redirect_stdout(io.StringIO()):
        mt = TransformerPoseEstimation(ModelConfig("transformer", **VIT_TRAIN_CFG))
    mt.load_state_dict(sd)
    mt.train()
    gt = torch.randn(2, 17, 3, generator=g) * 300
    pred = mt(img, dep, kp)
    total, comps = ComprehensivePoseLoss()(pred, gt)
    total.backward()
    names = [n for n, _ in mt.named_parameters()]
    gnorm = np.array([p.grad.double().norm().item() for _, p in mt.named_parameters()])
    # the restated functional oracle must give the same gradients (it is what the GPU tests differentiate)
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() and "grid" not in k else v) for k, v in sd.items()}
    po = tm.vit_forward(sdg, mt.config, img, dep, kp)
    to, _ = ComprehensivePoseLoss()(po, gt)
    to.backward()
    gn_o = np.array([sdg[n].grad.double().norm().item() for n in names])
    assert np.allclose(gn_o, gnorm, rtol=2e-3, atol=1e-6), np.abs(gn_o / np.maximum(gnorm, 1e-12) - 1).max()
    np.savez_compressed(os.path.join(OUT, "vit_256.npz"), versions=versions(), fill_seed=7, image_seed=9, kp=kp.numpy(),
                        out=out.numpy(), gt=gt.numpy(), train_pred=pred.detach().numpy(),
                        train_loss=np.float32(total.item()), grad_names=np.array(names), grad_norms=gnorm,
                        grad_final_cls=mt.final_cls_token.grad.numpy().reshape(-1),
                        grad_head_last_bias=mt.pose_head.decoder[-1].bias.grad.numpy(),
                        note="reference TransformerPoseEstimation over a restated timm ViT-B/16 stub; inputs torch.rand "
                             "with Generator(9): image, depth, kp, gt in that order")


CNN_TRAIN_CFG = dict(image_size=(256, 256), heatmap_size=256, regression_dropout=0.0)


def gen_cnn_train():
    """One training step of the live reference CNN (model.train(): batch-statistics BatchNorm; dropout 0 so that the step
    is deterministic) at 256x256, batch 4: predictions, loss, per-parameter gradient norms and the updated BatchNorm
    running statistics.  The functional oracle (oracle.torch_models.cnn_forward(train=True)) must reproduce them: it is
    what the GPU tests differentiate."""
    import torch
    from model_config import ModelConfig
    from models.cnn import CNNPoseEstimation
    from loss import ComprehensivePoseLoss
    sys.path.insert(0, REPO)
    from oracle import torch_models as tm
    cfg = ModelConfig("cnn", **CNN_TRAIN_CFG)
    m = CNNPoseEstimation(cfg)
    sd = tm.fill_state_dict(m.state_dict(), seed=5)
    m.load_state_dict(sd)
    m.train()
    g = torch.Generator().manual_seed(21)
    B = 4
    img = torch.rand(B, 3, 256, 256, generator=g)
    dep = torch.rand(B, 1, 256, 256, generator=g)
    kp = torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05
    kp[2, 7] = -1.0
    gt = torch.randn(B, 17, 3, generator=g) * 300
    pred = m(img, dep, kp)
    total, _ = ComprehensivePoseLoss()(pred, gt)
    total.backward()
    names = [n for n, _ in m.named_parameters()]
    gnorm = np.array([p.grad.double().norm().item() for _, p in m.named_parameters()])
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() and k in names else v.clone()) for k, v in sd.items()}
    po, new_stats = tm.cnn_forward(sdg, cfg, img, dep, kp, train=True, return_stats=True)
    to, _ = ComprehensivePoseLoss()(po, gt)
    to.backward()
    gn_o = np.array([sdg[n].grad.double().norm().item() for n in names])
    # parameters whose gradient is analytically zero (an affine shift that the next BatchNorm removes) hold rounding
    # noise six orders of magnitude below the real gradients: compare with an absolute floor
    assert np.allclose(gn_o, gnorm, rtol=5e-3, atol=2e-6 * gnorm.max()), np.abs(gn_o - gnorm).max()
    after = m.state_dict()
    for k, v in new_stats.items():
        assert torch.allclose(v, after[k], rtol=1e-4, atol=1e-5), k
    rm = np.array([after[k].double().sum().item() for k in sorted(after) if k.endswith("running_mean")])
    rv = np.array([after[k].double().sum().item() for k in sorted(after) if k.endswith("running_var")])
    np.savez_compressed(os.path.join(OUT, "cnn_train_256.npz"), versions=versions(), fill_seed=5, input_seed=21,
                        kp=kp.numpy(), gt=gt.numpy(), pred=pred.detach().numpy(), loss=np.float32(total.item()),
                        grad_names=np.array(names), grad_norms=gnorm, running_mean_sums=rm, running_var_sums=rv,
                        note="live reference CNNPoseEstimation.train() step, B=4, inputs torch.rand with Generator(21): "
                             "image, depth, kp, gt in that order")


def gen_metrics():
    """compute_mpjpe / compute_pa_mpjpe of the live reference (src/utils.py:55-165) on seeded poses incl. a mirrored pose
    (reflection branch), a pure similarity transform, identical poses and a collapsed prediction (variance <= 1e-9)."""
    import torch
    import utils
    rng = np.random.default_rng(17)
    B = 24
    gt = rng.normal(0, 300, (B, 17, 3)).astype(np.float32)
    pred = (gt + rng.normal(0, 60, (B, 17, 3))).astype(np.float32)
    pred[1] = gt[1] * np.array([-1, 1, 1], np.float32)                       # mirrored: det < 0 branch
    th = 0.7
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    pred[2] = (1.3 * gt[2] @ Rz + np.array([50, -20, 10], np.float32)).astype(np.float32)   # similarity transform
    pred[3] = gt[3]                                                           # identical
    pred[4] = np.float32(7.0)                                                 # collapsed: scale falls back to 1
    pred[5] = gt[5] @ Rz.T                                                    # pure rotation the other way
    tp, tg = torch.from_numpy(pred), torch.from_numpy(gt)
    per = np.array([utils.compute_pa_mpjpe(tp[i:i + 1], tg[i:i + 1]).item() for i in range(B)], np.float32)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), versions=versions(), pred=pred, gt=gt,
                        mpjpe=np.float32(utils.compute_mpjpe(tp, tg).item()),
                        pa_mpjpe=np.float32(utils.compute_pa_mpjpe(tp, tg).item()), pa_per_sample=per)


COLLATE_CASES = ([(20, 20), (31, 31), (26, 26), (23, 25)], [(16, 12)], [(8, 8), (8, 8)])


def collate_batch(sizes, seed):
    """The seeded sample list both this script and tests/test_gpu_parity.py build (CPU tensors)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    batch = []
    for i, (h, w) in enumerate(sizes):
        batch.append({"image": torch.rand(3, h, w, generator=g), "depth": torch.rand(1, h, w, generator=g),
                      "keypoints_2d": torch.rand(17, 2, generator=g), "joints_3d": torch.randn(17, 3, generator=g),
                      "camera_params": {"f": [1.0 + i, 1.0]}, "image_path": f"p{i}", "action": "a", "subaction": i % 2 + 1,
                      "image_size": torch.tensor([h, w]), "frame_idx": i})
    return batch


def gen_collate():
    """Human36MCollator of the live reference (src/dataset/collator.py:10-61) on ragged, single-sample and equal-size
    batches: padded images / depths, stacked key-points / joints / sizes and the padding field."""
    from dataset.collator import Human36MCollator
    col = Human36MCollator()
    out = {"versions": versions(), "n_cases": np.int32(len(COLLATE_CASES))}
    for c, sizes in enumerate(COLLATE_CASES):
        r = col(collate_batch(sizes, 900 + c))
        out[f"c{c}_sizes"] = np.array(sizes, np.int32)
        out[f"c{c}_image"] = r["image"].numpy()
        out[f"c{c}_depth"] = r["depth"].numpy()
        out[f"c{c}_keypoints_2d"] = r["keypoints_2d"].numpy()
        out[f"c{c}_joints_3d"] = r["joints_3d"].numpy()
        out[f"c{c}_image_size"] = r["image_size"].numpy()
        out[f"c{c}_padding"] = np.array(r["padding"], np.int32)
        out[f"c{c}_lists"] = json.dumps({k: r[k] for k in ("camera_params", "image_path", "action", "subaction", "frame_idx")})
    np.savez_compressed(os.path.join(OUT, "collate.npz"), **out)


RESIZE_CASES = ((3, 125, 126, 32, 32), (1, 100, 100, 64, 64), (3, 250, 251, 64, 64), (1, 37, 53, 16, 24), (3, 96, 128, 64, 64))


def resize_frame(case, seed):
    """Seeded uint8 frame (what torchvision.io.read_image returns) both this script and the tests build."""
    c, h, w = case[:3]
    return np.random.default_rng(seed).integers(0, 256, (c, h, w), dtype=np.uint8)


def gen_resize():
    """transforms.Resize(size) exactly as the reference applies it to a decoded frame (main.py:171-173,
    src/dataset/chunked_dataset.py:100-129): read_image -> .float() / 255.0 -> Resize; and the depth rescale of :159-164
    on the single-channel cases.  The C oracle's restatement of ATen's anti-aliased kernel must reproduce every frame bit
    for bit (asserted here, with the library versions recorded)."""
    import torch
    from torchvision import transforms
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    import oracle
    out = {"versions": versions(), "n_cases": np.int32(len(RESIZE_CASES)), "cases": np.array(RESIZE_CASES, np.int32)}
    for i, case in enumerate(RESIZE_CASES):
        u8 = resize_frame(case, 700 + i)
        x = torch.from_numpy(u8).float() / 255.0
        y = transforms.Compose([transforms.Resize((case[3], case[4]))])(x)
        assert np.array_equal(oracle.tensor_resize_aa(x.numpy(), case[3], case[4]), y.numpy()), case
        out[f"r{i}"] = y.numpy()
        if case[0] == 1:
            lo, hi = 0.35 + i, 7.25 + i
            out[f"r{i}_depth"] = (y * (hi - lo) + lo).numpy()
            out[f"r{i}_range"] = np.array([lo, hi], np.float64)
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, os.path.join(REF, "src"))
    os.chdir(tempfile.mkdtemp(prefix="pose_golden_"))
    which = sys.argv[1:] or ["augment", "heatmap", "loss", "config", "cnn", "vit", "cnn_train", "metrics", "collate", "resize"]
    for w in which:
        globals()["gen_" + w]()
        print("wrote", w)


if __name__ == "__main__":
    main()
