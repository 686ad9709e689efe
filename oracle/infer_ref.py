"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the inference input preparation (reference infer.py:217-221,
:362-367).  The arithmetic lives in a third-party dependency, torch (`F.interpolate(mode="bilinear",
align_corners=False)` = ATen upsample_bilinear2d, `area_pixel_compute_source_index`); torch is importable wherever the
tests run, so the restatement is pinned against the live call in tests/test_oracle_golden.py."""
import numpy as np


def depth_resize(depth, H, W):
    """depth [B, h, w] float32 -> [B, H, W]; ATen's source-index rule in float32, op by op."""
    depth = np.asarray(depth, dtype=np.float32)
    B, h, w = depth.shape
    f = np.float32

    def axis(n_in, n_out):
        scale = f(n_in) / f(n_out)
        src = np.maximum(scale * (np.arange(n_out, dtype=np.float32) + f(0.5)) - f(0.5), f(0))
        i0 = src.astype(np.int64)
        i1 = i0 + (i0 < n_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, (f(1) - l1).astype(np.float32), l1

    y0, y1, h0, h1 = axis(h, H)
    x0, x1, w0, w1 = axis(w, W)
    top = w0 * depth[:, y0][:, :, x0] + w1 * depth[:, y0][:, :, x1]
    bot = w0 * depth[:, y1][:, :, x0] + w1 * depth[:, y1][:, :, x1]
    return (h0[None, :, None] * top + h1[None, :, None] * bot).astype(np.float32)


def normalise_keypoints(kpts_px_conf, img_w, img_h):
    """[B, K, 3] (x_px, y_px, conf) -> ([B, K, 2], [B, K, 3]) with x / img_w, y / img_h (infer.py:217-221)."""
    k = np.asarray(kpts_px_conf, dtype=np.float32)
    out3 = k.copy()
    out3[..., 0] = k[..., 0] / np.float32(img_w)
    out3[..., 1] = k[..., 1] / np.float32(img_h)
    return out3[..., :2].copy(), out3
