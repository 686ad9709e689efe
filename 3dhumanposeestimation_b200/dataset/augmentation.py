"""PoseAugmentor backed by the fused sm_100a augmentation kernels.

Reference: src/dataset/augmentation.py:9-351.  Same constructor arguments and the same
``__call__(sample: dict) -> dict`` contract (sample schema: src/dataset/chunked_dataset.py:219-231).
The GPU path is batched: ``augment_batch`` takes stacked tensors and returns outputs zero-padded to a
common size the way the reference's collator pads variable-size samples (src/dataset/collator.py:20-44);
``__call__`` is the B = 1 wrapper around it.

Random numbers are drawn on the host from the global ``np.random`` state in the reference's order
(flip, angle, scale, tx, ty, brightness, contrast; a disabled stage draws nothing), or passed in
explicitly -- the kernels themselves are RNG-free.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..config import (BRIGHTNESS_RANGE, CONTRAST_RANGE, FLIP_PROB, ROTATION_RANGE, SCALE_RANGE, TRANSLATE_RANGE)

AUG_FLIP, AUG_ROTATE, AUG_SCALE, AUG_TRANSLATE, AUG_COLOR = 1, 2, 4, 8, 16


class PoseAugmentor:
    def __init__(
        self,
        rotation_range: Tuple[float, float] = ROTATION_RANGE,
        flip_prob: float = FLIP_PROB,
        scale_range: Tuple[float, float] = SCALE_RANGE,
        translate_range: Tuple[float, float] = TRANSLATE_RANGE,
        brightness_range: Tuple[float, float] = BRIGHTNESS_RANGE,
        contrast_range: Tuple[float, float] = CONTRAST_RANGE,
        enable_rotation: bool = True,
        enable_flip: bool = True,
        enable_scale: bool = True,
        enable_translate: bool = True,
        enable_color: bool = True,
    ):
        self.rotation_range = rotation_range
        self.flip_prob = flip_prob
        self.scale_range = scale_range
        self.translate_range = translate_range
        self.brightness_range = brightness_range
        self.contrast_range = contrast_range
        self.enable_rotation = enable_rotation
        self.enable_flip = enable_flip
        self.enable_scale = enable_scale
        self.enable_translate = enable_translate
        self.enable_color = enable_color
        self._ws = None
        self._plan_host = None
        self._plan_copied = None  # event: the previous async H2D copy of the pinned plan buffer

    # ------------------------------------------------------------------ host-side parameter draw
    @property
    def flags(self) -> int:
        return ((AUG_FLIP if self.enable_flip else 0) | (AUG_ROTATE if self.enable_rotation else 0)
                | (AUG_SCALE if self.enable_scale else 0) | (AUG_TRANSLATE if self.enable_translate else 0)
                | (AUG_COLOR if self.enable_color else 0))

    def draw_params(self, batch_size: int) -> np.ndarray:
        """[B, 8] fp64 {flip, angle, scale, tx_frac, ty_frac, brightness, contrast, 0}; per sample the
        global np.random stream is consumed exactly as augmentation.py:222-336 consumes it."""
        p = np.zeros((batch_size, 8), np.float64)
        for i in range(batch_size):
            if self.enable_flip:
                p[i, 0] = 1.0 if np.random.random() < self.flip_prob else 0.0
            if self.enable_rotation:
                p[i, 1] = np.random.uniform(self.rotation_range[0], self.rotation_range[1])
            if self.enable_scale:
                p[i, 2] = np.random.uniform(self.scale_range[0], self.scale_range[1])
            if self.enable_translate:
                p[i, 3] = np.random.uniform(self.translate_range[0], self.translate_range[1])
                p[i, 4] = np.random.uniform(self.translate_range[0], self.translate_range[1])
            if self.enable_color:
                p[i, 5] = np.random.uniform(self.brightness_range[0], self.brightness_range[1])
                p[i, 6] = np.random.uniform(self.contrast_range[0], self.contrast_range[1])
        return p

    # ------------------------------------------------------------------ batched GPU path
    def augment_batch(self, image: torch.Tensor, depth: torch.Tensor, keypoints_2d: torch.Tensor,
                      joints_3d: torch.Tensor, camera: torch.Tensor, params: Optional[np.ndarray] = None,
                      pad_to: Optional[Tuple[int, int]] = None) -> Dict[str, torch.Tensor]:
        """image [B,3,H,W] (fp32 in [0,1] or uint8), depth [B,1,H,W] (same dtype), keypoints_2d [B,J,2]
        fp32, joints_3d [B,J,3] fp32, camera [B,4] fp64 {fx,fy,cx,cy}; all CUDA.  Returns a dict with
        ``image`` [B,3,PH,PW] and ``depth`` [B,1,PH,PW] fp32 (sample i valid in the top-left
        ``sizes[i]`` corner, zero elsewhere), ``keypoints_2d``, ``joints_3d``, ``camera`` [B,4] fp64 and
        ``sizes`` [B,2] int32 (H', W')."""
        lib = _lib.lib()
        if image.dtype not in (torch.float32, torch.uint8) or depth.dtype != image.dtype:
            raise TypeError("image and depth must both be float32 or both be uint8")
        _lib.require_cuda(image, "image")
        _lib.require_cuda(depth, "depth")
        _lib.require_cuda(keypoints_2d, "keypoints_2d", torch.float32)
        _lib.require_cuda(joints_3d, "joints_3d", torch.float32)
        _lib.require_cuda(camera, "camera", torch.float64)
        if image.dim() != 4 or image.shape[1] != 3:
            raise ValueError(f"image must be [B,3,H,W], got {tuple(image.shape)}")
        B, _, H, W = image.shape
        if tuple(depth.shape) != (B, 1, H, W):
            raise ValueError(f"depth must be [B,1,H,W], got {tuple(depth.shape)}")
        J = keypoints_2d.shape[1]
        if tuple(keypoints_2d.shape) != (B, J, 2) or tuple(joints_3d.shape) != (B, J, 3):
            raise ValueError("keypoints_2d must be [B,J,2] and joints_3d [B,J,3]")
        if tuple(camera.shape) != (B, 4):
            raise ValueError("camera must be [B,4] = (fx, fy, cx, cy)")
        flags = self.flags
        if params is None:
            params = self.draw_params(B)
        params = np.ascontiguousarray(params, np.float64)
        if params.shape != (B, 8):
            raise ValueError(f"params must be [B, 8], got {params.shape}")

        # host: per-sample plan (libm / decimal rounding) + launch geometry
        nplan = B * _lib.POSE_AUG_PLAN_BYTES
        if self._plan_host is None or self._plan_host.numel() < nplan:
            self._plan_host = torch.empty(nplan, dtype=torch.uint8).pin_memory()
        capturing = torch.cuda.is_current_stream_capturing()
        if self._plan_copied is not None and not capturing:
            self._plan_copied.synchronize()  # the pinned buffer is about to be overwritten
        plan_host = self._plan_host[:nplan]
        launch = _lib.PoseAugLaunch()
        _lib.check(lib.pose_augment_plan(params.ctypes.data, B, H, W, flags, plan_host.data_ptr(), C.byref(launch)),
                   "pose_augment_plan")
        plan_dev = plan_host.to(image.device, non_blocking=True)
        if not capturing:   # (inside a CUDA-graph capture the copy is a graph node re-reading the pinned buffer)
            self._plan_copied = torch.cuda.Event()
            self._plan_copied.record()

        PH, PW = launch.max_out_h, (launch.max_out_w + 3) // 4 * 4
        if pad_to is not None:
            if pad_to[0] < launch.max_out_h or pad_to[1] < launch.max_out_w or pad_to[1] % 4:
                raise ValueError(f"pad_to={pad_to} must cover ({launch.max_out_h}, {launch.max_out_w}) with a "
                                 "width that is a multiple of 4")
            PH, PW = pad_to
        dev = image.device
        image_out = torch.empty((B, 3, PH, PW), dtype=torch.float32, device=dev)
        depth_out = torch.empty((B, 1, PH, PW), dtype=torch.float32, device=dev)
        kp_out = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
        joints_out = torch.empty((B, J, 3), dtype=torch.float32, device=dev)
        cam_out = torch.empty((B, 4), dtype=torch.float64, device=dev)
        sizes = torch.empty((B, 2), dtype=torch.int32, device=dev)
        nws = lib.pose_augment_workspace_bytes(B, H, W, C.byref(launch))
        if self._ws is None or self._ws.numel() < nws or self._ws.device != dev:
            self._ws = torch.zeros(nws, dtype=torch.uint8, device=dev)
        code = lib.pose_augment_batch(image.data_ptr(), depth.data_ptr(), 0 if image.dtype == torch.float32 else 1,
                                      keypoints_2d.data_ptr(), joints_3d.data_ptr(), camera.data_ptr(),
                                      plan_dev.data_ptr(), C.byref(launch), B, H, W, J, flags,
                                      image_out.data_ptr(), depth_out.data_ptr(), PH, PW, kp_out.data_ptr(),
                                      joints_out.data_ptr(), cam_out.data_ptr(), sizes.data_ptr(),
                                      self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr())
        _lib.check(code, "pose_augment_batch")
        return {"image": image_out, "depth": depth_out, "keypoints_2d": kp_out, "joints_3d": joints_out,
                "camera": cam_out, "sizes": sizes, "params": params}

    @staticmethod
    def launches_per_batch() -> int:
        """Kernels launched by one augment_batch call: operand pack, per-sample tables, fused cluster kernel."""
        return 3

    def kernel_error_flag(self) -> int:
        """Debug aid (synchronises): non-zero if the fused kernel ever saw a band that did not fit the
        launch geometry computed by pose_augment_plan."""
        if self._ws is None:
            return 0
        return int(self._ws[:4].view(torch.int32).item())

    # ------------------------------------------------------------------ reference-compatible call
    def __call__(self, sample: Dict) -> Dict:
        augmented = sample.copy()
        image, depth = sample["image"], sample["depth"]
        if not isinstance(image, torch.Tensor) or not isinstance(depth, torch.Tensor):
            raise TypeError("PoseAugmentor (B200) takes tensor samples (chunked_dataset.py:219-231 schema); "
                            "PIL inputs are not supported")
        src_device = image.device
        dev = src_device if src_device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        joints = torch.as_tensor(sample["joints_3d"], dtype=torch.float32)
        kp = torch.as_tensor(sample["keypoints_2d"], dtype=torch.float32)
        cam = sample["camera_params"]
        cam_t = torch.tensor([[cam["f"][0], cam["f"][1], cam["c"][0], cam["c"][1]]], dtype=torch.float64)
        out = self.augment_batch(image.to(dev).unsqueeze(0).contiguous(), depth.to(dev).unsqueeze(0).contiguous(),
                                 kp.to(dev).unsqueeze(0).contiguous(), joints.to(dev).unsqueeze(0).contiguous(),
                                 cam_t.to(dev))
        h, w = (int(v) for v in out["sizes"][0].tolist())
        augmented["image"] = out["image"][0, :, :h, :w].contiguous().to(src_device)
        augmented["depth"] = out["depth"][0, :, :h, :w].contiguous().to(src_device)
        augmented["joints_3d"] = out["joints_3d"][0].to(src_device)
        augmented["keypoints_2d"] = out["keypoints_2d"][0].to(src_device)
        if self.enable_scale:
            c = out["camera"][0].tolist()
            scaled = cam.copy()
            scaled["f"] = [c[0], c[1]]
            scaled["c"] = [c[2], c[3]]
            augmented["camera_params"] = scaled
        return augmented
