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


def split_k(tiles: int, kblocks: int, sms: int = 148, max_splits: int = 128) -> int:
    """Split-K factor of a weight-gradient contraction: `tiles` output tiles, `kblocks` 64-row blocks of contraction.
    Work items = tiles * splits run in rounds of one per SM; a count just above a multiple of the SM count wastes almost a
    whole round (measured: conv1.1's weight gradient 949 us at 300 items, 664 us at 295), so the factor minimises
    rounds x (blocks per item + a fixed per-item cost for the epilogue's atomics)."""
    best, best_cost = 1, None
    for s in range(1, max(1, min(max_splits, kblocks // 4)) + 1):
        rounds = -(-tiles * s // sms)
        cost = rounds * (-(-kblocks // s) + 8)
        if best_cost is None or cost < best_cost:
            best, best_cost = s, cost
    return best


def wgrad_splits(n_out: int, n_in: int, rows: int) -> int:
    """Split-K factor passed to pose_gemm_bf16_tr for dW[n_out, n_in] = dY^T X over `rows` samples.  split_k sizes it for
    128 x 128 tiles; when that lands on ONE split although the contraction is long and the output is 256-column tileable,
    two splits switch the kernel to 256-column CTA-pair tiles (the library doubles the factor for them) whose work items
    still fill the machine -- measured on the ViT's fc1 / fc2 weight gradients (768 x 3072 over 16448 rows): 90 -> 63 us."""
    kblocks = (rows + 63) // 64
    s = split_k(((n_out + 127) // 128) * ((n_in + 127) // 128), kblocks)
    if s == 1 and n_in % 256 == 0 and n_out >= 256 and kblocks >= 64:
        s = 2
    return s


class GraphedForward:
    """Eval-mode forward of a launch plan replayed from ONE CUDA graph (run_inference, infer.py:383-393, runs batch 1: a
    forward of ~200 launches of a few microseconds each is launch bound).  `fn(*static_inputs)` must enqueue the forward on
    the current stream and return a tensor that stays valid (a plan buffer).  Two eager calls warm the plan (allocations,
    kernel attributes), the third captures; `prepare()` runs before every replay for host-side work that must stay outside
    the graph (refreshing folded weights after a parameter update).  POSE_INFER_GRAPH=0 disables it."""

    MAX_BATCH = 8

    def __init__(self, fn, prepare=None):
        import os
        self.fn, self.prepare = fn, prepare
        self.calls, self.graph, self.static, self.out = 0, None, None, None
        self.enabled = os.environ.get("POSE_INFER_GRAPH", "1") != "0"

    def __call__(self, *inputs):
        import torch
        if not self.enabled or torch.cuda.is_current_stream_capturing():
            return self.fn(*inputs)
        self.calls += 1
        if self.calls <= 2:
            return self.fn(*inputs)
        if self.prepare is not None:
            self.prepare()
        if self.graph is None:
            self.static = [torch.empty_like(t) for t in inputs]
            for s_, t in zip(self.static, inputs):
                s_.copy_(t)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    self.out = self.fn(*self.static)
            except Exception:
                self.enabled = False
                torch.cuda.synchronize()
                return self.fn(*inputs)
            self.graph = graph
        else:
            for s_, t in zip(self.static, inputs):
                s_.copy_(t)
        self.graph.replay()
        return self.out
