"""Helpers of the reference that the hot path needs (reference: src/utils.py:55-69, :168-195)."""
import torch
import torch.nn as nn

# activation name -> id understood by the fused GEMM / elementwise epilogues (csrc)
ACT_NONE, ACT_RELU, ACT_SILU, ACT_GELU = 0, 1, 2, 3
_ACT_IDS = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "silu": ACT_SILU, "gelu": ACT_GELU}


def activation_id(name):
    """Fused-epilogue id of an activation name; the reference falls back to ReLU for unknown names
    (src/utils.py:180-181), and so does this."""
    if name in _ACT_IDS:
        return _ACT_IDS[name]
    if name in ("leaky_relu", "mish"):
        raise NotImplementedError(f"activation {name!r} has no sm_100a epilogue (BASELINE configs use silu / gelu)")
    return ACT_RELU


def get_activation(name):
    """Module factory kept for state-dict / repr compatibility (src/utils.py:168-181)."""
    if name == "relu":
        return nn.ReLU(inplace=True)
    if name == "silu":
        return nn.SiLU(inplace=True)
    if name == "gelu":
        return nn.GELU()
    if name == "leaky_relu":
        return nn.LeakyReLU(0.2, inplace=True)
    if name == "mish":
        return nn.Mish(inplace=True)
    return nn.ReLU(inplace=True)


def compute_mpjpe(predicted_joints, ground_truth_joints):
    """Mean per-joint position error (src/utils.py:55-69); plain tensor ops, used as the parity metric."""
    assert predicted_joints.shape == ground_truth_joints.shape
    errors = torch.linalg.norm(predicted_joints - ground_truth_joints, dim=2)
    return errors.mean(dim=1).mean()
