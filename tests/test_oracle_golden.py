"""CPU: the oracle restatement against goldens frozen from the LIVE reference (oracle/gen_golden.py).

These pin the oracle; the GPU parity tests then compare the CUDA path with the oracle."""
import json

import numpy as np


def _aug_cases(g):
    for i in range(int(g["n"])):
        yield i, f"c{i}_"


def test_augment_oracle_matches_reference_bit_exact(oracle, golden):
    g = golden("augment.npz")
    for i, k in _aug_cases(g):
        img = g[k + "image_u8"].astype(np.float32) / np.float32(255)
        dep = g[k + "depth_u8"].astype(np.float32) / np.float32(255)
        o = oracle.augment_sample(img, dep, g[k + "kp"], g[k + "joints"], g[k + "cam"], g[k + "params"],
                                  int(g[k + "flags"]))
        ref_img = g[k + "out_image_u8"].astype(np.float32) / np.float32(255)
        ref_dep = g[k + "out_depth_u8"].astype(np.float32) / np.float32(255)
        assert o["image"].shape == ref_img.shape, (i, o["image"].shape, ref_img.shape)
        assert np.array_equal(o["image"], ref_img), f"case {i}: RGB pixels differ"
        assert np.array_equal(o["depth"], ref_dep), f"case {i}: depth pixels differ"
        assert np.array_equal(o["keypoints_2d"].view(np.uint32), g[k + "out_kp"].view(np.uint32)), f"case {i}: kp"
        assert np.array_equal(o["joints_3d"].view(np.uint32), g[k + "out_joints"].view(np.uint32)), f"case {i}: joints"
        assert np.array_equal(o["cam"], g[k + "out_cam"]), f"case {i}: camera"


def test_heatmap_oracle_matches_reference(oracle, golden):
    g = golden("heatmap.npz")
    for name in ("s32", "vit", "cnn", "odd"):
        kp, hs, sigma = g[name + "_kp"], int(g[name + "_hs"]), float(g[name + "_sigma"])
        hm = oracle.heatmap(kp, hs, sigma)
        B = kp.shape[0]
        # arg-max location: bit-exact requirement
        assert np.array_equal(hm.reshape(B, 17, -1).argmax(-1), g[name + "_argmax"])
        peak = oracle.heatmap_peak(kp, hs)
        valid = peak >= 0
        assert np.array_equal(peak[valid], g[name + "_argmax"][valid])
        assert (g[name + "_max"][~valid] == 0).all() and (hm.reshape(B, 17, -1).max(-1)[~valid] == 0).all()
        # values: torch's vectorised exp and glibc expf differ by <= 1 ulp
        if name + "_full" in g:
            ref = g[name + "_full"]
            got = hm
        else:
            ref = g[name + "_rows"]
            got = hm[:, :, :: max(1, hs // 8), :]
        assert np.allclose(got, ref, rtol=2.5e-7, atol=1e-38)
        assert np.allclose(hm.astype(np.float64).sum((2, 3)), g[name + "_sum"], rtol=1e-6)


def test_loss_oracle_matches_reference(oracle, golden):
    g = golden("loss.npz")
    for name in ("b8", "b1", "b33_w"):
        out5, grad = oracle.pose_loss(g[name + "_pred"], g[name + "_gt"], g[name + "_weights"])
        assert np.allclose(out5, g[name + "_out5"], rtol=1e-5)  # north_star budget is 1e-3
        ref = g[name + "_grad"]
        assert np.abs(grad - ref).max() <= 1e-5 * np.abs(ref).max()
        assert np.all(np.isfinite(grad))


def test_oracle_stage_functions_edge_cases(oracle):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (9, 7, 3), dtype=np.uint8)
    # identity resize is a copy; identity affine (bilinear and fixed-point nearest) is a copy
    assert np.array_equal(oracle.resize_bilinear_aa(img, (9, 7)), img)
    ident = np.array([1, 0, 0, 0, 1, 0], np.float64)
    assert np.array_equal(oracle.affine_bilinear(img, ident), img)
    assert np.array_equal(oracle.affine_nearest_fixed(img, ident), img)
    # a shift by one whole image leaves only fill
    far = np.array([1, 0, 100, 0, 1, 0], np.float64)
    assert oracle.affine_bilinear(img, far).max() == 0
    assert oracle.scale_affine_nearest(img, (9, 7), far).max() == 0
    # brightness: truncation for factor <= 1, clipping above
    x = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    assert np.array_equal(oracle.brightness(x, 1.0), x)
    assert oracle.brightness(x, 1.9).max() == 255
    assert np.array_equal(oracle.brightness(x, 0.5)[..., 0].ravel(), (np.arange(256) * np.float32(0.5)).astype(np.uint8))
    # quantisation truncates
    assert np.array_equal(oracle.quantize_u8(np.array([0.0, 0.999, 1.0, 0.5], np.float32)), [0, 254, 255, 127])
    # rotate matrix: exact multiples of 90 take Pillow's transpose shortcuts
    assert oracle.rotate_matrix(0.0, 8, 8)[0] == 1 and oracle.rotate_matrix(360.0, 8, 8)[0] == 1
    assert oracle.rotate_matrix(180.0, 8, 6)[0] == 3 and oracle.rotate_matrix(-90.0, 8, 8)[0] == 4
    assert oracle.rotate_matrix(90.0, 8, 6)[0] == 0  # non-square: general affine path


def test_model_config_defaults_match_reference(golden):
    import importlib
    import os
    from conftest import GOLDEN
    mc = importlib.import_module("3dhumanposeestimation_b200.model_config")
    ref = json.load(open(os.path.join(GOLDEN, "model_config.json")))

    def norm(d):
        return json.loads(json.dumps(d))
    assert norm(mc.ModelConfig("cnn").to_dict()) == ref["cnn"]
    assert norm(mc.ModelConfig("transformer").to_dict()) == ref["transformer"]
    assert norm(mc.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256).to_dict()) == ref["cnn_256"]
    import pytest
    with pytest.raises(ValueError):
        mc.ModelConfig("rnn")


def test_metrics_oracle_matches_live_reference(oracle, golden):
    """compute_mpjpe / compute_pa_mpjpe (src/utils.py:55-165) incl. the mirrored, similarity, identical and collapsed poses."""
    d = golden("metrics.npz")
    assert abs(oracle.mpjpe(d["pred"], d["gt"]) - float(d["mpjpe"])) < 1e-3
    m, per = oracle.pa_mpjpe(d["pred"], d["gt"], per_sample=True)
    assert abs(m - float(d["pa_mpjpe"])) < 1e-3
    assert np.allclose(per, d["pa_per_sample"], rtol=1e-5, atol=1e-3)
    # the reference's rotation convention: a pose rotated about z by +0.7 rad (sample 2) is NOT perfectly re-aligned
    assert per[2] > 100.0 and per[3] < 1e-3


def test_infer_prep_oracle_matches_the_live_torch_call():
    """infer.py:362-367 calls F.interpolate(mode="bilinear", align_corners=False) directly: the numpy restatement is pinned
    against that live call (up- and down-sampling, non-square, odd sizes, 1-pixel axes)."""
    import torch
    import torch.nn.functional as F
    from oracle import infer_ref
    rng = np.random.default_rng(3)
    for (h, w, H, W) in ((37, 53, 256, 256), (480, 640, 256, 256), (256, 256, 256, 256), (1, 7, 5, 9), (300, 200, 512, 512)):
        d = rng.random((2, h, w), dtype=np.float32) * 10.0
        want = F.interpolate(torch.from_numpy(d)[:, None], size=(H, W), mode="bilinear", align_corners=False)[:, 0].numpy()
        got = infer_ref.depth_resize(d, H, W)
        assert np.allclose(got, want, rtol=2e-6, atol=2e-6), (h, w, H, W, np.abs(got - want).max())
    k = rng.random((2, 17, 3), dtype=np.float32) * np.float32(640.0)
    kp2, kp3 = infer_ref.normalise_keypoints(k, 640, 480)
    assert np.array_equal(kp2[..., 0], k[..., 0] / np.float32(640)) and np.array_equal(kp3[..., 2], k[..., 2])


def test_resize_oracle_reproduces_the_live_torchvision_resize(oracle, golden):
    """transforms.Resize as the reference applies it to decoded frames (main.py:171-173,
    chunked_dataset.py:100-129).  The C restatement of ATen's anti-aliased kernel equals the live outputs bit for bit
    (down-scaling with 5-9 taps, up-scaling, non-square frames)."""
    gold = golden("resize.npz")
    for i in range(int(gold["n_cases"])):
        c, h, w, oh, ow = (int(v) for v in gold["cases"][i])
        u8 = np.random.default_rng(700 + i).integers(0, 256, (c, h, w), dtype=np.uint8)
        x = u8.astype(np.float32) / np.float32(255.0)
        got = oracle.tensor_resize_aa(x, oh, ow)
        assert np.array_equal(got, gold[f"r{i}"]), (i, np.abs(got - gold[f"r{i}"]).max())
        if c == 1:
            lo, hi = (float(v) for v in gold[f"r{i}_range"])
            assert np.array_equal(got * np.float32(hi - lo) + np.float32(lo), gold[f"r{i}_depth"])
