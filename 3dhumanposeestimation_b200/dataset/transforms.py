"""``transforms.Resize(image_size)`` of decoded frames on the device.

Reference: ``transforms.Compose([transforms.Resize(model_cfg.image_size)])`` (main.py:171-173) applied to every decoded RGB
and depth frame in ``process_sample`` (src/dataset/chunked_dataset.py:100-129), then the depth rescale
``depth * (depth_max - depth_min) + depth_min`` (:159-164).  On a float tensor torchvision's ``Resize`` is
``F.interpolate(mode="bilinear", align_corners=False, antialias=True)``; ``pose_resize_bilinear_aa`` restates ATen's
separable anti-aliased resampling so that the result equals the reference's CPU tensor bit for bit (pinned by
``tests/golden/resize.npz``).  Frames may stay ``uint8`` (as ``torchvision.io.read_image`` returns them): the
``.float() / 255.0`` of the reference is applied on the fly.  CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _lib


class Resize:
    """Drop-in for ``transforms.Resize(size)`` on CHW / BCHW float32 (in [0, 1]) or uint8 CUDA tensors; bilinear,
    antialias=True (torchvision's default for tensors).  ``size``: int (both sides) or (h, w)."""

    def __init__(self, size):
        if isinstance(size, int):
            size = (size, size)
        size = tuple(int(s) for s in size)
        if len(size) != 2 or min(size) <= 0:
            raise ValueError("size must be an int or (h, w)")
        self.size = size

    def __call__(self, img: torch.Tensor) -> torch.Tensor:
        return resize_frames(img, self.size)

    def __repr__(self):
        return f"Resize(size={self.size}, interpolation=bilinear, antialias=True)"


def resize_frames(frames: torch.Tensor, size, depth_range=None) -> torch.Tensor:
    """frames [C, H, W] or [B, C, H, W], float32 in [0, 1] or uint8 -> float32 [.., size[0], size[1]].
    ``depth_range``: per-frame ``(depth_min, depth_max)`` pairs -- the reference's depth rescale fused into the same launch."""
    squeeze = frames.dim() == 3
    x = frames.unsqueeze(0) if squeeze else frames
    if x.dim() != 4:
        raise ValueError(f"expected [C, H, W] or [B, C, H, W], got {tuple(frames.shape)}")
    if x.dtype == torch.float32:
        in_dtype = 0
    elif x.dtype == torch.uint8:
        in_dtype = 1
    else:
        raise TypeError("frames must be float32 (in [0, 1]) or uint8")
    x = _lib.require_cuda(x.contiguous(), "frames")
    B, C, H, W = x.shape
    OH, OW = int(size[0]), int(size[1])
    lib = _lib.lib()
    nbytes = lib.pose_resize_workspace_bytes(H, W, OH, OW)
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=x.device)
    out = torch.empty((B, C, OH, OW), dtype=torch.float32, device=x.device)
    mul = add = None
    if depth_range is not None:
        if len(depth_range) != B:
            raise ValueError("one (depth_min, depth_max) pair per frame")
        # python floats, cast to fp32 exactly as `tensor * (max - min) + min` does
        mul = torch.tensor([float(hi) - float(lo) for lo, hi in depth_range], dtype=torch.float32).to(x.device)
        add = torch.tensor([float(lo) for lo, _ in depth_range], dtype=torch.float32).to(x.device)
    code = lib.pose_resize_bilinear_aa(x.data_ptr(), in_dtype, B, C, H, W, OH, OW, mul.data_ptr() if mul is not None else None,
                                       add.data_ptr() if add is not None else None, ws.data_ptr(), nbytes, out.data_ptr(),
                                       _lib.stream_ptr())
    _lib.check(code, "pose_resize_bilinear_aa")
    return out[0] if squeeze else out
