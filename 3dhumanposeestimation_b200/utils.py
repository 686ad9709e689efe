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


def eval_metrics(predicted_joints, ground_truth_joints):
    """(means [2], per_sample [2, B]) fp32 device tensors: row 0 MPJPE, row 1 PA-MPJPE -- one launch for both metrics
    (src/utils.py:55-165; the reference loops over samples in Python with a 3x3 SVD each)."""
    from . import _lib
    if predicted_joints.shape != ground_truth_joints.shape:
        raise AssertionError(f"Shape mismatch: pred {predicted_joints.shape}, gt {ground_truth_joints.shape}")
    pred = _lib.require_cuda(predicted_joints.detach().float().contiguous(), "predicted_joints", torch.float32)
    gt = _lib.require_cuda(ground_truth_joints.detach().float().contiguous(), "ground_truth_joints", torch.float32)
    if pred.dim() != 3 or pred.shape[2] != 3:
        raise ValueError(f"expected [N, num_joints, 3], got {tuple(pred.shape)}")
    B, J = pred.shape[0], pred.shape[1]
    per = torch.empty(2, B, dtype=torch.float32, device=pred.device)
    means = torch.empty(2, dtype=torch.float32, device=pred.device)
    _lib.check(_lib.lib().pose_eval_metrics(pred.data_ptr(), gt.data_ptr(), B, J, per.data_ptr(), means.data_ptr(),
                                            _lib.stream_ptr()), "pose_eval_metrics")
    return means, per


def compute_mpjpe(predicted_joints, ground_truth_joints):
    """Mean per-joint position error (src/utils.py:55-69): 0-dim device tensor."""
    return eval_metrics(predicted_joints, ground_truth_joints)[0][0]


def compute_pa_mpjpe(predicted_joints, ground_truth_joints):
    """Procrustes-aligned MPJPE (src/utils.py:72-165, including its rotation convention): 0-dim device tensor."""
    return eval_metrics(predicted_joints, ground_truth_joints)[0][1]
