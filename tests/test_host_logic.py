"""CPU: host-side logic of the drop-in classes (no kernels launched)."""
import importlib

import numpy as np
import pytest


def test_draw_params_consumes_rng_in_reference_order(pose, golden):
    """The goldens record the explicit draws of the live reference for each seed/constructor."""
    import json
    g = golden("augment.npz")
    for i in range(int(g["n"])):
        k = f"c{i}_"
        aug = pose.PoseAugmentor(**json.loads(str(g[k + "ctor"])))
        assert aug.flags == int(g[k + "flags"])
        np.random.seed(int(g[k + "seed"]))
        p = aug.draw_params(1)[0]
        assert np.array_equal(p, g[k + "params"]), (i, p, g[k + "params"])


def test_draw_params_batch_is_sequential(pose):
    aug = pose.PoseAugmentor()
    np.random.seed(5)
    a = aug.draw_params(3)
    np.random.seed(5)
    b = np.stack([aug.draw_params(1)[0] for _ in range(3)])
    assert np.array_equal(a, b)


def test_state_dict_keys_of_common_modules(pose):
    hm = pose.GaussianHeatmapGenerator(17, 64, 2.0)
    assert sorted(hm.state_dict()) == ["x_grid", "y_grid"]
    assert hm.state_dict()["x_grid"][3, 5] == 5 and hm.state_dict()["y_grid"][3, 5] == 3
    head = pose.PoseRegressionHead(1024, 17, hidden_dims=[1024, 512], dropout=0.2, activation="silu")
    assert sorted(head.state_dict()) == sorted(
        ["decoder.0.0.weight", "decoder.0.0.bias", "decoder.1.0.weight", "decoder.1.0.bias", "decoder.2.weight",
         "decoder.2.bias"])
    assert head.state_dict()["decoder.2.weight"].shape == (51, 512)


def test_loss_module_attributes(pose):
    crit = pose.ComprehensivePoseLoss()
    assert (crit.mse_weight, crit.l1_weight, crit.inter_joint_loss_weight, crit.abs_root_loss_weight) == (1, 1, 100, 1)
    crit = pose.ComprehensivePoseLoss(l1_weight=0.5, mse_weight=2.0, inter_joint_loss_weight=10.0,
                                      abs_root_loss_weight=3.0)
    assert crit._weights() == (2.0, 0.5, 10.0, 3.0)


def test_cnn_state_dict_layout_matches_reference(pose):
    """Checkpoint compatibility: same keys and shapes as the reference's CNNPoseEstimation."""
    import json
    import os
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "cnn_state_dict_layout.json")))
    m = pose.CNNPoseEstimation(pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == ref
    assert sum(p.numel() for p in m.parameters()) == 26920792
    assert m.config.to_dict()["stage_channels"] == [128, 256, 512]


def test_torch_oracle_cnn_matches_reference_golden(golden):
    """The fp32 model oracle (oracle/torch_models.py) reproduces the live reference on the small configuration."""
    import json
    import torch
    from oracle import torch_models as tm
    import importlib
    pose = importlib.import_module("3dhumanposeestimation_b200")
    g = golden("cnn_small.npz")
    cfg = pose.ModelConfig("cnn", **json.loads(str(g["config"])))
    m = pose.CNNPoseEstimation(cfg)
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    sd = tm.fill_state_dict(m.state_dict(), seed=int(g["fill_seed"]))
    out = tm.cnn_forward(sd, cfg, torch.from_numpy(g["image"]), torch.from_numpy(g["depth"]), torch.from_numpy(g["kp"]))
    assert np.abs(out.numpy() - g["out"]).max() < 1e-3 * max(1.0, np.abs(g["out"]).max())


def test_checkpoint_round_trip_in_the_reference_format(tmp_path):
    """src/train.py:300-309 layout out, infer.py:73-131 semantics in (bare state dict, 'module.' prefixes, model_args)."""
    import importlib
    import torch
    pose = importlib.import_module("3dhumanposeestimation_b200")
    ck = importlib.import_module("3dhumanposeestimation_b200.checkpoint")
    cfg = pose.ModelConfig("cnn", image_size=(64, 64), heatmap_size=64, initial_channels=32, stage_channels=[64, 128, 256],
                           global_pool_size=2, global_feature_dim=256, regression_dims=[128, 64])
    m = pose.CNNPoseEstimation(cfg)
    path = str(tmp_path / "ckpt_cnn_step_5.pth")
    ck.save_checkpoint(path, m, "cnn", step=5)
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"step", "model_state_dict", "optimizer_state_dict", "model_args", "model_type"}
    assert raw["model_type"] == "cnn" and raw["model_args"]["stage_channels"] == [64, 128, 256]
    m2 = ck.load_pose_model(path, "transformer", device="cpu")          # model_type comes from the file
    assert type(m2).__name__ == "CNNPoseEstimation" and not m2.training
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert set(sd) == set(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    # DataParallel-style "module." prefixes are stripped (infer.py:95-98)
    raw["model_state_dict"] = {"module." + k: v for k, v in sd.items()}
    torch.save(raw, path)
    m3 = ck.load_pose_model(path, "cnn", device="cpu")
    assert all(torch.equal(sd[k], v) for k, v in m3.state_dict().items())
    # a file that is neither a state dict nor a checkpoint dictionary is rejected like the reference does
    torch.save([1, 2, 3], path)
    try:
        ck.load_pose_model(path, "cnn", device="cpu")
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def test_next_row_host_mirrors_refuse_cpu_tensors():
    """infer.prepare_model_inputs / Human36MCollator: no CPU fallback (the product path fails loudly without the GPU)."""
    import importlib
    import pytest
    import torch
    infer = importlib.import_module("3dhumanposeestimation_b200.infer")
    col = importlib.import_module("3dhumanposeestimation_b200.dataset.collator").Human36MCollator()
    with pytest.raises(Exception):
        infer.prepare_model_inputs(torch.rand(1, 1, 8, 8), torch.rand(1, 17, 3), (8, 8), (16, 16))
    sample = {"image": torch.rand(3, 8, 8), "depth": torch.rand(1, 8, 8), "keypoints_2d": torch.rand(17, 2),
              "joints_3d": torch.rand(17, 3), "camera_params": {}, "image_path": "p", "action": "a", "subaction": 1,
              "image_size": torch.tensor([8, 8]), "frame_idx": 0}
    with pytest.raises(Exception):
        col([sample])
    with pytest.raises(ValueError):
        col([])


def test_vit_pretrained_weights_from_a_local_timm_state_dict(pose, tmp_path, monkeypatch):
    """ModelConfig("transformer") defaults to vit_pretrained=True (reference model_config.py / transformers.py:174-214): the
    timm weights come from POSE_VIT_WEIGHTS; the 3 -> 4 channel patch embedding is adapted like the reference (RGB filters
    kept, the depth channel gets their mean) and the position embedding is resampled from the 14 x 14 pre-training grid."""
    import torch
    tr = importlib.import_module("3dhumanposeestimation_b200.models.transformers")
    torch.manual_seed(3)
    donor = tr.VisionTransformerBackbone("vit_base_patch16_224", (224, 224), 3)
    sd = {k: torch.randn_like(v) * 0.05 for k, v in donor.state_dict().items()}
    sd["head.weight"], sd["head.bias"] = torch.zeros(1000, 768), torch.zeros(1000)     # classifier of the checkpoint: dropped
    path = tmp_path / "vit_base_patch16_224.pth"
    torch.save(sd, path)
    monkeypatch.setenv("POSE_VIT_WEIGHTS", str(path))
    cfg = pose.ModelConfig("transformer", image_size=(256, 256))
    assert cfg.vit_pretrained is True
    m = pose.TransformerPoseEstimation(cfg)
    got = m.vit_backbone.state_dict()
    w = got["patch_embed.proj.weight"]
    assert tuple(w.shape) == (768, 4, 16, 16)
    assert torch.equal(w[:, :3], sd["patch_embed.proj.weight"])
    assert torch.allclose(w[:, 3:], sd["patch_embed.proj.weight"].mean(dim=1, keepdim=True))
    pe = got["pos_embed"]
    assert tuple(pe.shape) == (1, 257, 768) and torch.equal(pe[:, 0], sd["pos_embed"][:, 0])
    ref = torch.nn.functional.interpolate(sd["pos_embed"][:, 1:].reshape(1, 14, 14, 768).permute(0, 3, 1, 2), size=(16, 16),
                                          mode="bicubic", antialias=True).permute(0, 2, 3, 1).reshape(1, 256, 768)
    assert torch.allclose(pe[:, 1:], ref)
    assert torch.equal(got["blocks.5.attn.qkv.weight"], sd["blocks.5.attn.qkv.weight"])
    monkeypatch.delenv("POSE_VIT_WEIGHTS")
    with pytest.raises(NotImplementedError, match="POSE_VIT_WEIGHTS"):
        pose.TransformerPoseEstimation(pose.ModelConfig("transformer", image_size=(256, 256)))


def test_weight_gradient_split_factors(pose):
    """utils.wgrad_splits: wave-aware split-K for 128 x 128 tiles, and two splits (-> 256-column CTA-pair tiles inside the
    library) where that tiling would be a single wave over a long contraction (the ViT's fc1 / fc2 weight gradients)."""
    u = importlib.import_module("3dhumanposeestimation_b200.utils")
    rows = 64 * 257
    assert u.wgrad_splits(768, 3072, rows) == 2 and u.wgrad_splits(3072, 768, rows) == 2      # 144 tiles: one wave -> 2
    assert u.wgrad_splits(2304, 768, rows) == 4 and u.wgrad_splits(768, 768, rows) == 4        # already split: unchanged
    assert u.wgrad_splits(768, 3072, 64 * 16) == u.split_k(144, 16)                            # short contraction: unchanged
    assert u.wgrad_splits(51, 512, rows) == u.split_k(4, 257)                                   # narrow output: unchanged
    for tiles, kb in ((1, 4), (36, 2048), (300, 37), (148, 257)):
        s = u.split_k(tiles, kb)
        assert 1 <= s <= max(1, kb // 4)
