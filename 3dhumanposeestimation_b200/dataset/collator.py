"""Human36MCollator (reference: src/dataset/collator.py:10-61) on the device.

Same call signature and result dictionary: a list of sample dictionaries in, the collated batch out -- images and depth
maps zero padded on the right / bottom to the largest sample of the batch and stacked, key-points / joints / image sizes
stacked, the remaining fields gathered into lists, ``"padding": [(max_h, max_w)] * B``.  The padding + stacking of all
samples is ONE kernel launch over a device table of sample pointers (the reference runs 2 B ``F.pad`` calls and two
``torch.stack``).  The sample tensors must live on the GPU (what ``PoseAugmentor`` returns here); CPU tensors raise --
there is no CPU fallback.  `depth_range` optionally applies the dataset's depth rescale
``depth * (max - min) + min`` (src/dataset/chunked_dataset.py:159-164) on the way.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

from .. import _lib


class Human36MCollator:
    def __call__(self, batch, depth_range=None):
        if not batch:
            raise ValueError("empty batch")
        B = len(batch)
        imgs, deps = [], []
        for s in batch:
            img = _lib.require_cuda(s["image"].contiguous(), "image", torch.float32)
            dep = _lib.require_cuda(s["depth"].contiguous(), "depth", torch.float32)
            if img.dim() != 3 or img.shape[0] != 3 or dep.dim() != 3 or dep.shape[0] != 1 or dep.shape[1:] != img.shape[1:]:
                raise ValueError(f"sample shapes: image {tuple(img.shape)}, depth {tuple(dep.shape)}")
            imgs.append(img)
            deps.append(dep)
        max_height = max(t.shape[1] for t in imgs)
        max_width = max(t.shape[2] for t in imgs)
        dev = imgs[0].device
        raw = bytearray()
        for i, (img, dep) in enumerate(zip(imgs, deps)):
            lo, hi = (0.0, 1.0) if depth_range is None else (float(depth_range[i][0]), float(depth_range[i][1]))
            raw += struct.pack("<QQiiff", img.data_ptr(), dep.data_ptr(), img.shape[1], img.shape[2], hi - lo, lo)
        table = torch.from_numpy(np.frombuffer(bytes(raw), dtype=np.uint8).copy()).to(dev)
        image = torch.empty(B, 3, max_height, max_width, dtype=torch.float32, device=dev)
        depth = torch.empty(B, 1, max_height, max_width, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().pose_collate_pad(table.data_ptr(), B, max_height, max_width, image.data_ptr(), depth.data_ptr(),
                                               _lib.stream_ptr()), "pose_collate_pad")
        return {
            "image": image,
            "depth": depth,
            "keypoints_2d": torch.stack([s["keypoints_2d"] for s in batch]),
            "joints_3d": torch.stack([s["joints_3d"] for s in batch]),
            "camera_params": [s["camera_params"] for s in batch],
            "image_path": [s["image_path"] for s in batch],
            "action": [s["action"] for s in batch],
            "subaction": [s["subaction"] for s in batch],
            "image_size": torch.stack([s["image_size"] for s in batch]),
            "frame_idx": [s["frame_idx"] for s in batch],
            "padding": [(max_height, max_width)] * B,
        }
