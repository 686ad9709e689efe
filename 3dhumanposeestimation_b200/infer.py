"""Inference path of the reference (infer.py) on the device: model-input preparation and `run_inference`.

The third-party detectors of the reference (YOLO pose: infer.py:134-237, DepthPro: infer.py:240-252) are out of scope
(pretrained models, no network); what follows them is reproduced:

* `prepare_model_inputs`: the depth map is resized to the model input size with
  ``F.interpolate(mode="bilinear", align_corners=False)`` (infer.py:362-367) and the detected pixel key-points are
  normalised by the image size (infer.py:217-221) -- one C-ABI call, `pose_infer_prep`;
* `run_inference` (infer.py:383-393): eval-mode forward under ``no_grad``, first sample of the batch as numpy.
"""
from __future__ import annotations

import torch

from . import _lib


def prepare_model_inputs(depth_map: torch.Tensor, keypoints_px_conf: torch.Tensor | None, image_size_wh, model_input_size):
    """depth_map [B, 1, h, w] (or [B, h, w]) fp32 CUDA; keypoints_px_conf [B, K, 3] = (x_pixel, y_pixel, confidence) or None;
    image_size_wh = (img_w, img_h) of the original image; model_input_size = (H, W).

    Returns (transformed_depth [B, 1, H, W], keypoints_2d_for_model [B, K, 2], keypoints_2d_with_conf [B, K, 3]) --
    the tensors `preprocess_input` hands to the model and to the visualisation (infer.py:370-376)."""
    d = _lib.require_cuda(depth_map.detach().float().contiguous(), "depth_map", torch.float32)
    if d.dim() == 4:
        if d.shape[1] != 1:
            raise ValueError(f"depth_map: expected one channel, got {tuple(d.shape)}")
        d = d[:, 0]
    if d.dim() != 3:
        raise ValueError(f"depth_map: expected [B, 1, h, w] or [B, h, w], got {tuple(depth_map.shape)}")
    B, h, w = d.shape
    H, W = int(model_input_size[0]), int(model_input_size[1])
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=d.device)
    kp2 = kp3 = None
    kptr = kp2ptr = kp3ptr = None
    K = 0
    img_w, img_h = float(image_size_wh[0]), float(image_size_wh[1])
    if keypoints_px_conf is not None:
        k = _lib.require_cuda(keypoints_px_conf.detach().float().contiguous(), "keypoints_px_conf", torch.float32)
        if k.dim() != 3 or k.shape[0] != B or k.shape[2] != 3:
            raise ValueError(f"keypoints_px_conf: expected [{B}, K, 3], got {tuple(k.shape)}")
        K = k.shape[1]
        kp2 = torch.empty(B, K, 2, dtype=torch.float32, device=d.device)
        kp3 = torch.empty(B, K, 3, dtype=torch.float32, device=d.device)
        kptr, kp2ptr, kp3ptr = k.data_ptr(), kp2.data_ptr(), kp3.data_ptr()
    _lib.check(_lib.lib().pose_infer_prep(d.data_ptr(), B, h, w, H, W, out.data_ptr(), kptr, K, img_w, img_h, kp2ptr, kp3ptr,
                                          _lib.stream_ptr()), "pose_infer_prep")
    return out, kp2, kp3


def run_inference(pose_model, image_tensor, depth_tensor, keypoints_2d_tensor):
    """infer.py:383-393: eval-mode forward under no_grad; returns the first sample's joints [J, 3] as numpy."""
    pose_model.eval()
    with torch.no_grad():
        predicted = pose_model(image_tensor, depth_tensor, keypoints_2d_tensor)
    return predicted[0].float().cpu().numpy()
