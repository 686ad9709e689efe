"""GaussianHeatmapGenerator and PoseRegressionHead (reference: src/models/common.py:6-89).

Same constructor arguments, forward signatures and state-dict keys (``x_grid`` / ``y_grid`` buffers;
``decoder.{i}.0.{weight,bias}`` for hidden layers, ``decoder.{n}.{weight,bias}`` for the output
layer); the bodies run hand-written sm_100a kernels through the C ABI.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from ..utils import get_activation


def render_heatmaps(keypoints_2d, heatmap_size, sigma, out=None, dtype=torch.float32, channels_last=False,
                    c_stride=None, c_offset=0):
    """pose_heatmap_render.  Default: [B, J, hs, hs] fp32 planes exactly like the reference.  With
    ``channels_last`` the planes are written as channels ``c_offset .. c_offset+J`` of an existing
    [B, hs, hs, c_stride] tensor (the CNN's conv1 operand) so they never exist as separate planes."""
    kp = _lib.require_cuda(keypoints_2d, "keypoints_2d", torch.float32)
    if kp.dim() != 3 or kp.shape[2] != 2:
        raise ValueError(f"keypoints_2d must be [B, J, 2], got {tuple(kp.shape)}")
    B, J = kp.shape[0], kp.shape[1]
    hs = int(heatmap_size)
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("heat-maps are rendered in fp32 or bf16")
    if channels_last:
        if out is None:
            raise ValueError("channels_last rendering writes into an existing [B, hs, hs, C] tensor")
        _lib.require_cuda(out, "out", dtype)
        c_stride = out.shape[-1]
        if tuple(out.shape) != (B, hs, hs, c_stride):
            raise ValueError(f"out must be [B, hs, hs, C], got {tuple(out.shape)}")
    else:
        if out is None:
            out = torch.empty((B, J, hs, hs), dtype=dtype, device=kp.device)
        _lib.require_cuda(out, "out", dtype)
        c_stride = 0
    code = _lib.lib().pose_heatmap_render(kp.data_ptr(), B, J, hs, float(sigma), out.data_ptr(),
                                          0 if dtype == torch.float32 else 1, 1 if channels_last else 0,
                                          int(c_stride), int(c_offset), _lib.stream_ptr())
    _lib.check(code, "pose_heatmap_render")
    return out


class GaussianHeatmapGenerator(nn.Module):
    def __init__(self, num_joints, heatmap_size=64, sigma=2.0):
        super().__init__()
        self.num_joints = num_joints
        self.heatmap_size = heatmap_size
        self.sigma = sigma
        # kept only because they are part of the reference's state_dict (common.py:18-21); the kernel
        # generates coordinates from thread indices
        coords = torch.arange(heatmap_size, dtype=torch.float32)
        y_grid, x_grid = torch.meshgrid(coords, coords, indexing="ij")
        self.register_buffer("x_grid", x_grid.clone())
        self.register_buffer("y_grid", y_grid.clone())

    def forward(self, keypoints_2d):
        kp = keypoints_2d.detach()
        if kp.dtype != torch.float32:
            kp = kp.float()
        return render_heatmaps(kp.contiguous(), self.heatmap_size, self.sigma)


class PoseRegressionHead(nn.Module):
    def __init__(self, in_features, num_joints, hidden_dims=(512, 256), dropout=0.2, activation="gelu"):
        super().__init__()
        self.num_joints = num_joints
        self.activation = activation
        self.dropout = dropout
        layers = []
        prev_dim = in_features
        for hidden_dim in hidden_dims:
            layers.append(nn.Sequential(nn.Linear(prev_dim, hidden_dim), get_activation(activation),
                                        nn.Dropout(dropout)))
            prev_dim = hidden_dim
        layers.append(nn.Linear(prev_dim, num_joints * 3))
        self.decoder = nn.Sequential(*layers)

    def forward(self, x):
        from ..ops import mlp_head_forward  # tcgen05 GEMM path
        x = x.reshape(x.size(0), -1)
        linears = [m[0] if isinstance(m, nn.Sequential) else m for m in self.decoder]
        pose = mlp_head_forward(x, linears, self.activation, self.dropout if self.training else 0.0)
        return pose.view(-1, self.num_joints, 3)
