"""CNNPoseEstimation on B200 (reference: src/models/cnn.py:9-665).

The module tree (class names, attribute names, parameter shapes) reproduces the reference's so that
``state_dict()`` keys match and reference checkpoints load unchanged (src/train.py:300-309,
infer.py:73-131).  The modules are parameter containers: ``forward`` does not walk them with PyTorch
ops but runs a flat launch plan of hand-written sm_100a kernels over channels-last bf16 activations
(``CnnInferencePlan``): BatchNorm folded into the convolution weights, dense convolutions as implicit
GEMM on tcgen05 with bias / SiLU / residual / concat fused in the epilogue, depthwise convolutions
fused with their BatchNorm, activation and SE/ECA squeeze, heat-maps rendered straight into the
21-channel conv1 operand.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _lib
from ..utils import GraphedForward, activation_id, get_activation
from .common import GaussianHeatmapGenerator, PoseRegressionHead


def _norm(name, channels):
    if name != "batch":
        raise NotImplementedError(f"normalization {name!r}: the B200 path implements BatchNorm2d (reference default)")
    return nn.BatchNorm2d(channels)


# ------------------------------------------------------------------------------------------------------
# parameter containers (same attribute names as the reference => same state_dict keys)
# ------------------------------------------------------------------------------------------------------
class SEBlock(nn.Module):  # cnn.py:9-26
    def __init__(self, channels, reduction=16, activation="silu"):
        super().__init__()
        self.activation = activation
        self.fc = nn.Sequential(nn.Linear(channels, channels // reduction, bias=False), get_activation(activation),
                                nn.Linear(channels // reduction, channels, bias=False), nn.Sigmoid())


class ECABlock(nn.Module):  # cnn.py:29-45
    def __init__(self, channels, gamma=2, b=1):
        super().__init__()
        t = int(abs(math.log(channels, 2) + b) / gamma)
        k = t if t % 2 else t + 1
        self.conv = nn.Conv1d(1, 1, kernel_size=k, padding=(k - 1) // 2, bias=False)


class CoordAttention(nn.Module):  # cnn.py:48-98
    def __init__(self, in_channels, out_channels, reduction=32):
        super().__init__()
        mid = max(8, in_channels // reduction)
        self.conv1 = nn.Conv2d(in_channels, mid, 1)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv_h = nn.Conv2d(mid, out_channels, 1)
        self.conv_w = nn.Conv2d(mid, out_channels, 1)


class ConvBnAct(nn.Module):  # cnn.py:101-139
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=None, groups=1, bias=False,
                 activation="silu", normalization="batch", dilation=1):
        super().__init__()
        if padding is None:
            padding = (kernel_size - 1) // 2 * dilation
        if bias:
            raise NotImplementedError("ConvBnAct with a conv bias is not used by the reference model")
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, groups=groups, bias=False,
                              dilation=dilation)
        self.norm = _norm(normalization, out_channels)
        self.activation = activation  # name or None


class DepthwiseSeparableConv(nn.Module):  # cnn.py:142-186
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, activation="silu", normalization="batch"):
        super().__init__()
        self.depthwise = ConvBnAct(in_channels, in_channels, kernel_size, stride, groups=in_channels,
                                   activation=activation, normalization=normalization)
        self.pointwise = ConvBnAct(in_channels, out_channels, 1, activation=activation, normalization=normalization)


def _attention(kind, channels, se_reduction, activation):
    if kind == "se":
        return SEBlock(channels, se_reduction, activation)
    if kind == "eca":
        return ECABlock(channels)
    if kind == "coord":
        return CoordAttention(channels, channels)
    return None


class InvertedResidual(nn.Module):  # cnn.py:189-266
    def __init__(self, in_channels, out_channels, stride=1, expand_ratio=6, use_se=True, se_reduction=16,
                 activation="silu", normalization="batch", residual_scale=1.0, attention_type=None):
        super().__init__()
        self.stride = stride
        self.use_residual = in_channels == out_channels and stride == 1
        self.residual_scale = residual_scale
        hidden = int(in_channels * expand_ratio)
        layers = []
        if expand_ratio != 1:
            layers.append(ConvBnAct(in_channels, hidden, 1, activation=activation, normalization=normalization))
        layers.append(ConvBnAct(hidden, hidden, stride=stride, groups=hidden, activation=activation,
                                normalization=normalization))
        kind = "se" if (attention_type == "se" or (use_se and attention_type is None)) else attention_type
        att = _attention(kind, hidden, se_reduction, activation)
        if att is not None:
            layers.append(att)
        layers.append(ConvBnAct(hidden, out_channels, 1, activation=None, normalization=normalization))
        self.conv = nn.Sequential(*layers)


class DualPathBlock(nn.Module):  # cnn.py:269-380
    def __init__(self, in_channels, out_channels, stride=1, activation="silu", normalization="batch",
                 residual_scale=1.0, attention_type=None):
        super().__init__()
        self.residual_scale = residual_scale
        self.stride = stride
        kw = dict(activation=activation, normalization=normalization)
        self.residual_path = nn.Sequential(
            ConvBnAct(in_channels, out_channels, 1, **kw),
            DepthwiseSeparableConv(out_channels, out_channels, stride=stride, **kw),
            ConvBnAct(out_channels, out_channels, 1, activation=None, normalization=normalization))
        dense = out_channels // 2
        self.dense_path = nn.Sequential(ConvBnAct(in_channels, dense, 1, **kw),
                                        DepthwiseSeparableConv(dense, dense, stride=stride, **kw))
        self.attention = _attention(attention_type, out_channels, 16, activation)
        self.fusion = ConvBnAct(out_channels + dense, out_channels, 1, **kw)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_channels != out_channels:
            self.shortcut = ConvBnAct(in_channels, out_channels, 1, stride=stride, activation=None,
                                      normalization=normalization)


class WASPModule(nn.Module):  # cnn.py:383-479
    def __init__(self, in_channels, out_channels, dilations=(1, 6, 12, 18), activation="silu", normalization="batch"):
        super().__init__()
        kw = dict(activation=activation, normalization=normalization)
        self.dilations = tuple(dilations)
        self.conv1x1 = ConvBnAct(in_channels, out_channels, 1, **kw)
        self.atrous_branches = nn.ModuleList(
            [ConvBnAct(in_channels, out_channels, 3, padding=d, dilation=d, **kw) for d in dilations])
        self.global_branch = nn.Sequential(nn.AdaptiveAvgPool2d(1), ConvBnAct(in_channels, out_channels, 1, **kw))
        n = len(dilations) + 2
        self.weights = nn.Parameter(torch.ones(n) / n)
        self.fusion = ConvBnAct(out_channels, out_channels, 1, **kw)


class CNNPoseEstimation(nn.Module):  # cnn.py:482-665
    def __init__(self, config):
        super().__init__()
        self.config = config
        c = config
        kw = dict(activation=c.activation, normalization=c.normalization)
        self.conv1 = nn.Sequential(
            ConvBnAct(c.in_channels, c.initial_channels, c.initial_kernel_size, c.initial_stride, **kw),
            ConvBnAct(c.initial_channels, c.initial_channels, 3, 1, **kw))
        self.heatmap_generator = GaussianHeatmapGenerator(c.num_joints, c.heatmap_size, c.heatmap_sigma)
        self.stages = nn.ModuleList()
        cin = c.initial_channels
        for i, cout in enumerate(c.stage_channels):
            blocks = []
            for j in range(c.stage_depths[i]):
                stride = c.stage_strides[i] if j == 0 else 1
                src = cin if j == 0 else cout
                dual = i >= 2 and c.use_dual_path_blocks and (j == 0 or j % 2 == 0)
                if dual:
                    blocks.append(DualPathBlock(src, cout, stride, residual_scale=c.residual_scale,
                                                attention_type="coord" if i >= 2 else "se", **kw))
                else:
                    att = ("coord" if i >= 2 else "se") if j == 0 else ("eca" if j % 2 == 0 else "se")
                    blocks.append(InvertedResidual(src, cout, stride, c.stage_expand_ratios[i], c.use_se_blocks,
                                                   c.se_reduction, residual_scale=c.residual_scale,
                                                   attention_type=att, **kw))
            self.stages.append(nn.Sequential(*blocks))
            cin = cout
        last = c.stage_channels[-1]
        self.wasp = WASPModule(last, last, (1, 6, 12, 18), **kw)
        self.global_features = nn.Sequential(nn.AdaptiveAvgPool2d(c.global_pool_size),
                                             ConvBnAct(last, c.global_feature_dim, 1, **kw),
                                             ECABlock(c.global_feature_dim), nn.AdaptiveAvgPool2d(1))
        self.pose_head = PoseRegressionHead(c.global_feature_dim, c.num_joints, hidden_dims=c.regression_dims,
                                            dropout=c.regression_dropout, activation=c.activation)
        self._initialize_weights()
        self._plans = {}

    def _initialize_weights(self):  # cnn.py:627-639
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, image, depth, keypoints_2d):
        """image [B,3,S,S], depth [B,1,S,S], keypoints_2d [B,J,2] (fp32, CUDA) -> joints [B,J,3] fp32."""
        if self.training and torch.is_grad_enabled():
            plan = self.plan(image.shape[0], image.device)
            return _CnnTrainFn.apply(plan, image, depth, keypoints_2d, self.wasp.weights)
        if self.training:
            raise NotImplementedError("training-mode forward under no_grad (batch statistics without backward): "
                                      "call model.eval() for inference")
        B = image.shape[0]
        key = (B, image.device.index)
        plan = self._plans.get(key)
        if plan is None:
            plan = CnnInferencePlan(self, B, image.device)
            self._plans[key] = plan
        return plan.run(image, depth, keypoints_2d)

    def plan(self, B, device):
        """Training launch plan (forward with batch statistics + backward) at batch size B."""
        from .cnn_train import CnnTrainPlan
        key = ("train", B, device.index)
        p = self._plans.get(key)
        if p is None or not p.flat.intact():
            p = CnnTrainPlan(self, B, device)
            self._plans[key] = p
        return p


class _CnnTrainFn(torch.autograd.Function):
    """Autograd node of the whole model: backward runs the plan's reverse pass, which accumulates straight into the
    parameters' flat ``.grad`` views (the `anchor` parameter only ties the node into the graph)."""

    @staticmethod
    def forward(ctx, plan, image, depth, kp, anchor):
        ctx.plan = plan
        out = plan.forward(image, depth, kp, save=True)
        return out.view(-1, plan.J, 3).clone()

    @staticmethod
    def backward(ctx, dout):
        ctx.plan.backward(dout.contiguous().view(dout.shape[0], -1))
        return None, None, None, None, None


# ------------------------------------------------------------------------------------------------------
# inference launch plan
# ------------------------------------------------------------------------------------------------------
def _fold_bn(conv_w, bn):
    """eval-mode BatchNorm folded into the preceding bias-free convolution (fp32): W' = W * g / sqrt(v + eps),
    b' = beta - mean * g / sqrt(v + eps)."""
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv_w.detach().float() * s.view(-1, *([1] * (conv_w.dim() - 1)))
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * s
    return w, b


class CnnInferencePlan:
    """Flat list of kernel launches for eval-mode CNNPoseEstimation.forward at a fixed batch size.
    Buffers are allocated once; folded bf16 weights are rebuilt whenever a parameter or buffer of the
    model changes (in-place torch writes bump tensor versions; the fused AdamW bumps FlatParams.generation)."""

    def __init__(self, model: CNNPoseEstimation, B: int, device):
        self.model, self.B, self.dev = model, B, device
        self.lib = _lib.lib()
        c = model.config
        S = int(c.heatmap_size)
        if tuple(c.image_size) != (S, S):
            raise ValueError("CNNPoseEstimation needs heatmap_size == image height == width (cnn.py:648)")
        if c.in_channels != 4 + c.num_joints or c.in_channels > 32:
            raise NotImplementedError("conv1 operand packs 3 RGB + 1 depth + J heat-maps into 32 channels")
        self.S, self.J = S, c.num_joints
        self.act = activation_id(c.activation)
        self.steps = []      # closures
        self.weights = []    # (rebuild_fn)
        self.keep = []       # tensors that must stay alive
        self._version = None
        self._tracked = None
        self._graphed = None
        self.last_epi = None
        self.launches = 0
        self._build()

    # ---- helpers -------------------------------------------------------------------------------------
    def _buf(self, *shape, dtype=torch.bfloat16, zero=False):
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.dev)
        self.keep.append(t)
        return t

    def _parts(self, npx):
        """Pixel chunks per image for squeeze sums: enough CTAs to fill the chip, at least 8 pixels per chunk."""
        return max(1, min(npx // 8 if npx >= 8 else 1, (148 * 4 + self.B - 1) // self.B))

    def _param_version(self):
        if self._tracked is None:
            self._tracked = list(self.model.parameters()) + list(self.model.buffers())
        # pose_adamw_step writes through raw pointers (no _version bump): FlatParams.generation records those writes
        ent = getattr(self._tracked[0], "_pose_flat", None)
        gen = ent[0].generation if ent is not None else 0
        return sum(t._version for t in self._tracked) + sum(t.data_ptr() & 0xFFFF for t in self._tracked[:4]) + gen

    def _epi(self, out, ldc, bias, act, out_scale=1.0, residual=None, ldr=0, res_scale=0.0, fp32=False, col_off=0):
        e = _lib.PoseGemmEpilogue()
        e.bias = bias.data_ptr() if bias is not None else None
        e.residual = residual.data_ptr() if residual is not None else None
        e.C = out.data_ptr() + col_off * (4 if fp32 else 2)
        e.ldc, e.ldr, e.act, e.out_dtype = ldc, ldr, act, 0 if fp32 else 1
        e.out_scale, e.res_scale = float(out_scale), float(res_scale)
        self.keep.append(e)
        self.last_epi = e
        return e

    def _launch(self, fn, *args):
        lib, name = self.lib, fn

        def step():
            _lib.check(getattr(lib, name)(*args, _lib.stream_ptr()), name)
        self.steps.append(step)
        self.launches += 1

    def _dense_weight(self, cba: ConvBnAct, cin_pad=None, scale_param=None):
        """KRSC bf16 weight [Cout, KH*KW*Cin_pad] + fp32 bias, refreshed by self.weights closures."""
        conv = cba.conv
        cout, cin, kh, kw = conv.weight.shape
        cin_pad = cin_pad or cin
        w16 = self._buf(cout, kh * kw * cin_pad)
        b32 = self._buf(cout, dtype=torch.float32)

        def rebuild():
            w, b = _fold_bn(conv.weight, cba.norm)
            w = w.permute(0, 2, 3, 1)  # KRSC
            if cin_pad != cin:
                w = torch.nn.functional.pad(w, (0, cin_pad - cin))
            w16.copy_(w.reshape(cout, -1))
            b32.copy_(b)
        self.weights.append(rebuild)
        return w16, b32

    # ---- layer emitters ------------------------------------------------------------------------------
    def conv1x1(self, x, cba, out=None, ldc=None, col_off=0, act="default", residual=None, out_scale=1.0,
                res_scale=0.0):
        """x [B,H,W,Cin] -> [B,H,W,Cout] (or a column slice of `out`)."""
        Bn, H, W, cin = x.shape
        cout = cba.conv.out_channels
        w16, b32 = self._dense_weight(cba)
        if out is None:
            out = self._buf(Bn, H, W, cout)
        ldc = ldc or out.shape[-1]
        a = 0 if (act is None or cba.activation is None) else (self.act if act == "default" else activation_id(act))
        e = self._epi(out, ldc, b32, a, out_scale, residual, residual.shape[-1] if residual is not None else 0,
                      res_scale, col_off=col_off)
        self._launch("pose_gemm_bf16_ex", x.data_ptr(), cin, w16.data_ptr(), cin, Bn * H * W, cout, cin, C.byref(e))
        return out

    def conv2d(self, x, cba, act="default", out=None, residual=None, out_scale=1.0, res_scale=0.0, cin_pad=None):
        Bn, H, W, cin_x = x.shape
        conv = cba.conv
        cout, _, kh, kw = conv.weight.shape
        stride, dil, pad = conv.stride[0], conv.dilation[0], conv.padding[0]
        w16, b32 = self._dense_weight(cba, cin_pad=cin_x)
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        if out is None:
            out = self._buf(Bn, Ho, Wo, cout)
        a = 0 if (act is None or cba.activation is None) else (self.act if act == "default" else activation_id(act))
        e = self._epi(out, out.shape[-1], b32, a, out_scale, residual,
                      residual.shape[-1] if residual is not None else 0, res_scale)
        if kh == kw and kh > 1 and stride == 1 and 2 * pad == dil * (kh - 1) and dil >= H and dil >= W:
            # the dilation reaches across the whole map (WASP rate 18 on a 16 x 16 map): every tap but the centre reads only
            # zero padding, so the layer is the 1x1 convolution with the centre-tap weights (KRSC rows of pitch K*K*Cin)
            ctr = (kh // 2 * kw + kw // 2) * cin_x
            self._launch("pose_gemm_bf16_ex", x.data_ptr(), cin_x, w16.data_ptr() + 2 * ctr, kh * kw * cin_x, Bn * H * W, cout,
                         cin_x, C.byref(e))
            return out
        self._launch("pose_conv2d_bf16", x.data_ptr(), Bn, H, W, cin_x, w16.data_ptr(), cout, kh, kw, stride, dil, pad,
                     C.byref(e))
        return out

    def dwconv(self, x, cba, want_pool=False):
        Bn, H, W, ch = x.shape
        conv = cba.conv
        if conv.kernel_size != (3, 3) or conv.padding != (1, 1) or conv.groups != ch:
            raise NotImplementedError("depthwise kernel: 3x3, padding 1")
        stride = conv.stride[0]
        wd = self._buf(9, ch, dtype=torch.float32)
        b32 = self._buf(ch, dtype=torch.float32)

        def rebuild():
            w, b = _fold_bn(conv.weight, cba.norm)       # [C,1,3,3]
            wd.copy_(w.reshape(ch, 9).t())
            b32.copy_(b)
        self.weights.append(rebuild)
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        out = self._buf(Bn, Ho, Wo, ch)
        a = self.act if cba.activation is not None else 0
        pool = self._buf(Bn, self.lib.pose_dwconv3x3_pool_parts(H, W, stride), ch, dtype=torch.float32) if want_pool else None
        self._launch("pose_dwconv3x3_bf16", x.data_ptr(), Bn, H, W, ch, wd.data_ptr(), b32.data_ptr(), stride, a,
                     out.data_ptr(), pool.data_ptr() if pool is not None else None,
                     pool.shape[1] if pool is not None else 0)
        return (out, pool) if want_pool else out

    def pool_sums(self, x):
        """[B, parts, C] fp32 partial channel sums of x [B,H,W,C]."""
        Bn, H, W, ch = x.shape
        pool = self._buf(Bn, self._parts(H * W), ch, dtype=torch.float32)
        self._launch("pose_pool_sum_bf16", x.data_ptr(), Bn, H * W, ch, pool.data_ptr(), pool.shape[1])
        return pool

    def attention(self, x, att, pool=None):
        """SE / ECA / CoordAttention on x [B,H,W,C]; `pool` = channel sums of x if the producer made them."""
        Bn, H, W, ch = x.shape
        if isinstance(att, (SEBlock, ECABlock)):
            if pool is None:
                pool = self.pool_sums(x)
            parts = pool.shape[1]
            gate = self._buf(Bn, ch, dtype=torch.float32)
            if isinstance(att, SEBlock):
                # both Linear layers as tcgen05 GEMMs over the whole batch (M = B): the hidden width is zero-padded
                # to one 64-wide K block
                w1, w2 = att.fc[0].weight, att.fc[2].weight
                cr = w1.shape[0]
                crp = (cr + 63) // 64 * 64
                w1p, w2p = self._buf(crp, ch, zero=True), self._buf(ch, crp, zero=True)
                self.weights.append(lambda: (w1p[:cr].copy_(w1.detach()), w2p[:, :cr].copy_(w2.detach())))
                mean16 = self._buf(Bn, ch)
                self._launch("pose_sums_to_bf16", pool.data_ptr(), parts, Bn, ch, 1.0 / (H * W), mean16.data_ptr())
                hid = self._buf(Bn, crp)
                e1 = self._epi(hid, crp, None, activation_id(att.activation))
                self._launch("pose_gemm_bf16_ex", mean16.data_ptr(), ch, w1p.data_ptr(), ch, Bn, crp, ch, C.byref(e1))
                e2 = self._epi(gate, ch, None, 4, fp32=True)
                self._launch("pose_gemm_bf16_ex", hid.data_ptr(), crp, w2p.data_ptr(), crp, Bn, ch, crp, C.byref(e2))
            else:
                k = att.conv.weight.shape[-1]
                wk = self._buf(k, dtype=torch.float32)
                self.weights.append(lambda: wk.copy_(att.conv.weight.detach().view(-1)))
                self._launch("pose_eca_gate", pool.data_ptr(), parts, 1.0 / (H * W), wk.data_ptr(), k, Bn, ch,
                             gate.data_ptr(), None)
            self._launch("pose_channel_affine_bf16", x.data_ptr(), gate.data_ptr(), None, Bn, H * W, ch, x.data_ptr())
            return x
        if isinstance(att, CoordAttention):
            mid = att.conv1.out_channels
            midp = 64                                        # pad the tiny mid width to one K block
            P = self._buf(Bn, H + W, ch)
            self._launch("pose_coord_pool_bf16", x.data_ptr(), Bn, H, W, ch, P.data_ptr())
            w1 = self._buf(midp, ch, zero=True)
            b1 = self._buf(midp, dtype=torch.float32, zero=True)
            w2 = self._buf(2 * ch, midp, zero=True)
            b2 = self._buf(2 * ch, dtype=torch.float32)

            def rebuild():
                s = att.bn1.weight.detach() / torch.sqrt(att.bn1.running_var.detach() + att.bn1.eps)
                w1[:mid].copy_(att.conv1.weight.detach().view(mid, ch) * s[:, None])
                b1[:mid].copy_((att.conv1.bias.detach() - att.bn1.running_mean.detach()) * s + att.bn1.bias.detach())
                w2[:ch, :mid].copy_(att.conv_h.weight.detach().view(ch, mid))
                w2[ch:, :mid].copy_(att.conv_w.weight.detach().view(ch, mid))
                b2[:ch].copy_(att.conv_h.bias.detach())
                b2[ch:].copy_(att.conv_w.bias.detach())
            self.weights.append(rebuild)
            Y = self._buf(Bn * (H + W), midp)
            e1 = self._epi(Y, midp, b1, activation_id("silu"))
            self._launch("pose_gemm_bf16_ex", P.data_ptr(), ch, w1.data_ptr(), ch, Bn * (H + W), midp, ch, C.byref(e1))
            G = self._buf(Bn, H + W, 2 * ch)
            e2 = self._epi(G, 2 * ch, b2, 4)
            self._launch("pose_gemm_bf16_ex", Y.data_ptr(), midp, w2.data_ptr(), midp, Bn * (H + W), 2 * ch, midp,
                         C.byref(e2))
            out = self._buf(Bn, H, W, ch)
            self._launch("pose_coord_apply_bf16", x.data_ptr(), G.data_ptr(), Bn, H, W, ch, out.data_ptr())
            return out
        raise NotImplementedError(type(att))

    def inverted_residual(self, x, blk: InvertedResidual):
        cbas = [mod for mod in blk.conv if isinstance(mod, ConvBnAct)]
        atts = [mod for mod in blk.conv if not isinstance(mod, ConvBnAct)]
        y = x
        if len(cbas) == 3:                        # 1x1 expansion only exists for expand_ratio != 1
            y = self.conv1x1(y, cbas[0])
        dw, proj = cbas[-2], cbas[-1]
        att = atts[0] if atts else None
        pool = None
        if isinstance(att, (SEBlock, ECABlock)):
            y, pool = self.dwconv(y, dw, want_pool=True)
        else:
            y = self.dwconv(y, dw)
        if att is not None:
            y = self.attention(y, att, pool)
        if blk.use_residual:   # x + conv(x) * residual_scale
            return self.conv1x1(y, proj, act=None, residual=x, out_scale=blk.residual_scale, res_scale=1.0)
        return self.conv1x1(y, proj, act=None)

    def dual_path(self, x, blk: DualPathBlock):
        Bn, H, W, _ = x.shape
        cout = blk.fusion.conv.out_channels
        dense = blk.dense_path[0].conv.out_channels
        s = blk.stride
        r = self.conv1x1(x, blk.residual_path[0])
        r = self.dwconv(r, blk.residual_path[1].depthwise)
        r = self.conv1x1(r, blk.residual_path[1].pointwise)
        if isinstance(blk.shortcut, ConvBnAct):
            sc = self.conv2d(x, blk.shortcut, act=None) if s != 1 else self.conv1x1(x, blk.shortcut, act=None)
        else:
            sc = x
        Ho, Wo = r.shape[1], r.shape[2]
        cat = self._buf(Bn, Ho, Wo, cout + dense)        # torch.cat([res_path, dense_path], 1) by column slices
        self.conv1x1(r, blk.residual_path[2], out=cat, ldc=cout + dense, col_off=0, act=None, residual=sc,
                     out_scale=1.0, res_scale=blk.residual_scale)
        d = self.conv1x1(x, blk.dense_path[0])
        d = self.dwconv(d, blk.dense_path[1].depthwise)
        self.conv1x1(d, blk.dense_path[1].pointwise, out=cat, ldc=cout + dense, col_off=cout)
        out = self.conv1x1(cat, blk.fusion)
        if blk.attention is not None:
            out = self.attention(out, blk.attention)
        return out

    def wasp(self, x, m: WASPModule):
        Bn, H, W, ch = x.shape
        # the softmax-ed branch weights are runtime values: the epilogue scales are patched whenever the
        # weights are (re)built
        scales = []
        acc = self._buf(Bn, H, W, m.conv1x1.conv.out_channels)
        self.conv1x1(x, m.conv1x1, out=acc)
        scales.append((self.last_epi, 0))
        for i, br in enumerate(m.atrous_branches):
            self.conv2d(x, br, out=acc, residual=acc, res_scale=1.0)
            scales.append((self.last_epi, i + 1))
        pool = self.pool_sums(x)
        mean16 = self._buf(Bn, ch)
        self._launch("pose_sums_to_bf16", pool.data_ptr(), pool.shape[1], Bn, ch, 1.0 / (H * W), mean16.data_ptr())
        g = self.conv1x1(mean16.view(Bn, 1, 1, ch), m.global_branch[1])
        scales.append((self.last_epi, len(m.dilations) + 1))
        # bilinear interpolation of a 1x1 map is a broadcast (cnn.py:465-467)
        self._launch("pose_channel_affine_bf16", acc.data_ptr(), None, g.data_ptr(), Bn, H * W, acc.shape[-1],
                     acc.data_ptr())

        def set_scales():
            w = torch.softmax(m.weights.detach().float(), 0).cpu().tolist()
            for e, idx in scales:
                e.out_scale = float(w[idx])
        self.weights.append(set_scales)
        return self.conv1x1(acc, m.fusion)

    # ---- whole model ---------------------------------------------------------------------------------
    def _build(self):
        m, Bn, S = self.model, self.B, self.S
        c = m.config
        self.x0 = self._buf(Bn, S, S, 32)
        self.inputs = None   # set per call
        x = self.conv2d(self.x0, m.conv1[0])
        x = self.conv2d(x, m.conv1[1])
        for stage in m.stages:
            for blk in stage:
                x = self.dual_path(x, blk) if isinstance(blk, DualPathBlock) else self.inverted_residual(x, blk)
        x = self.wasp(x, m.wasp)
        gp = int(c.global_pool_size)
        H = x.shape[1]
        if H == 2 * gp:
            y = self._buf(Bn, gp, gp, x.shape[-1])
            self._launch("pose_avgpool2x2_bf16", x.data_ptr(), Bn, H, H, x.shape[-1], y.data_ptr())
            x = y
        elif H > gp:
            y = self._buf(Bn, gp, gp, x.shape[-1])
            self._launch("pose_adaptive_avgpool_bf16", x.data_ptr(), Bn, H, H, x.shape[-1], gp, gp, y.data_ptr())
            x = y
        elif H != gp:
            raise NotImplementedError(f"AdaptiveAvgPool2d({gp}) from a smaller {H}x{H} map (up-sampling) is not built")
        x = self.conv1x1(x, m.global_features[1])
        ch = x.shape[-1]
        pool = self.pool_sums(x)
        eca = m.global_features[2]
        k = eca.conv.weight.shape[-1]
        wk = self._buf(k, dtype=torch.float32)
        self.weights.append(lambda: wk.copy_(eca.conv.weight.detach().view(-1)))
        feat = self._buf(Bn, ch)
        self._launch("pose_eca_gate", pool.data_ptr(), pool.shape[1], 1.0 / (gp * gp), wk.data_ptr(), k, Bn, ch, None,
                     feat.data_ptr())
        # regression head: Linear + act on the tcgen05 GEMM (dropout is the identity in eval mode)
        lins = [mod[0] if isinstance(mod, nn.Sequential) else mod for mod in m.pose_head.decoder]
        h = feat
        for i, lin in enumerate(lins):
            last = i == len(lins) - 1
            n_out, n_in = lin.weight.shape
            w16 = self._buf(n_out, n_in)
            b32 = self._buf(n_out, dtype=torch.float32)
            self.weights.append(lambda w16=w16, b32=b32, lin=lin: (w16.copy_(lin.weight.detach()),
                                                                   b32.copy_(lin.bias.detach())))
            out = self._buf(Bn, n_out, dtype=torch.float32 if last else torch.bfloat16)
            e = self._epi(out, n_out, b32, 0 if last else self.act, fp32=last)
            self._launch("pose_gemm_bf16_ex", h.data_ptr(), n_in, w16.data_ptr(), n_in, Bn, n_out, n_in, C.byref(e))
            h = out
        self.out = h

    def run(self, image, depth, kp):
        _lib.require_cuda(image, "image", torch.float32)
        _lib.require_cuda(depth, "depth", torch.float32)
        _lib.require_cuda(kp, "keypoints_2d", torch.float32)
        Bn, S = self.B, self.S
        if tuple(image.shape) != (Bn, 3, S, S) or tuple(depth.shape) != (Bn, 1, S, S) or tuple(kp.shape) != (Bn, self.J, 2):
            raise ValueError(f"expected image [{Bn},3,{S},{S}], depth [{Bn},1,{S},{S}], keypoints [{Bn},{self.J},2]")
        if Bn <= GraphedForward.MAX_BATCH:
            # small batches are launch bound: the forward is replayed from one CUDA graph (weights are refreshed in place,
            # outside the graph, whenever a parameter changed)
            if self._graphed is None:
                self._graphed = GraphedForward(self._enqueue, prepare=self._refresh_weights)
            return self._graphed(image, depth, kp).view(Bn, self.J, 3).clone()
        return self._enqueue(image, depth, kp).view(Bn, self.J, 3).clone()

    def _refresh_weights(self):
        v = self._param_version()
        if v != self._version:
            with torch.no_grad():
                for rebuild in self.weights:
                    rebuild()
            self._version = v

    def _enqueue(self, image, depth, kp):
        Bn, S = self.B, self.S
        if not torch.cuda.is_current_stream_capturing():
            self._refresh_weights()
        _lib.check(self.lib.pose_cnn_input_pack(image.data_ptr(), depth.data_ptr(), kp.data_ptr(), Bn, S, self.J,
                                                float(self.model.config.heatmap_sigma), self.x0.data_ptr(),
                                                _lib.stream_ptr()), "pose_cnn_input_pack")
        for step in self.steps:
            step()
        return self.out
