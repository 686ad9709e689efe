"""Plain-PyTorch fp32 restatement of the reference's model forwards -- TEST INFRASTRUCTURE ONLY.

``cnn_forward(sd, cfg, image, depth, kp, train)`` evaluates CNNPoseEstimation.forward
(reference: src/models/cnn.py:641-665 and the block classes :9-479) as a pure function of a
``state_dict`` with the reference's key names, using nothing but ``torch.nn.functional`` calls.
It is the floating-point oracle for the bf16 tensor-core path (tolerance 0.5 mm MPJPE,
BASELINE.json:north_star) and is itself pinned against the live reference by
``tests/golden/cnn_small.npz`` (oracle/gen_golden.py).

Nothing in the product package imports this module.
"""
from __future__ import annotations

import math
import zlib

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _act(x, name):
    if name is None:
        return x
    if name == "silu":
        return F.silu(x)
    if name == "gelu":
        return F.gelu(x)
    return F.relu(x)


class _Ctx:
    """Walks a state_dict by prefix; in train mode batch statistics are used and the updated running
    statistics are collected in ``new_stats`` (momentum 0.1, unbiased variance -- nn.BatchNorm2d)."""

    def __init__(self, sd, train, act, trace=None):
        self.sd, self.train, self.act, self.new_stats, self.trace = sd, train, act, {}, trace

    def bn(self, x, p):
        w, b = self.sd[p + ".weight"], self.sd[p + ".bias"]
        rm, rv = self.sd[p + ".running_mean"], self.sd[p + ".running_var"]
        if self.train:
            rm, rv = rm.clone(), rv.clone()
            y = F.batch_norm(x, rm, rv, w, b, True, 0.1, BN_EPS)
            self.new_stats[p + ".running_mean"], self.new_stats[p + ".running_var"] = rm, rv
            return y
        return F.batch_norm(x, rm, rv, w, b, False, 0.1, BN_EPS)

    def conv_bn_act(self, x, p, stride=1, act="default", dilation=1, groups=1):
        """ConvBnAct (cnn.py:101-139): conv without bias, 'same'-style padding, BN, activation."""
        w = self.sd[p + ".conv.weight"]
        k = w.shape[-1]
        pad = (k - 1) // 2 * dilation
        x = F.conv2d(x, w, None, stride, pad, dilation, groups)
        x = self.bn(x, p + ".norm")
        x = _act(x, self.act if act == "default" else act)
        if self.trace is not None:
            self.trace[p] = x.detach()
        return x

    def dw(self, x, p, stride=1):
        return self.conv_bn_act(x, p, stride=stride, groups=x.shape[1])

    def se(self, x, p):  # cnn.py:9-26
        y = x.mean((2, 3))
        y = _act(F.linear(y, self.sd[p + ".fc.0.weight"]), self.act)
        y = torch.sigmoid(F.linear(y, self.sd[p + ".fc.2.weight"]))
        return x * y[:, :, None, None]

    def eca(self, x, p):  # cnn.py:29-45
        w = self.sd[p + ".conv.weight"]
        y = x.mean((2, 3))[:, None, :]
        y = torch.sigmoid(F.conv1d(y, w, padding=(w.shape[-1] - 1) // 2))[:, 0]
        return x * y[:, :, None, None]

    def coord(self, x, p):  # cnn.py:48-98
        n, c, h, w = x.shape
        xh = x.mean(3, keepdim=True)                     # [n,c,h,1]
        xw = x.mean(2, keepdim=True)                     # [n,c,1,w]
        cat = torch.cat([xh.transpose(2, 3), xw], 3)     # [n,c,1,h+w]
        y = F.conv2d(cat, self.sd[p + ".conv1.weight"], self.sd[p + ".conv1.bias"])
        y = F.silu(self.bn(y, p + ".bn1"))
        yh, yw = y[..., :h].transpose(2, 3), y[..., h:]
        ah = torch.sigmoid(F.conv2d(yh, self.sd[p + ".conv_h.weight"], self.sd[p + ".conv_h.bias"]))
        aw = torch.sigmoid(F.conv2d(yw, self.sd[p + ".conv_w.weight"], self.sd[p + ".conv_w.bias"]))
        return x * ah * aw

    def attention(self, x, p):
        if p + ".fc.0.weight" in self.sd:
            return self.se(x, p)
        if p + ".conv1.weight" in self.sd:
            return self.coord(x, p)
        return self.eca(x, p)


def _inverted_residual(c, x, p, stride, expand, residual_scale):  # cnn.py:189-266
    sd = c.sd
    i = 0
    y = x
    if expand != 1:                              # 1x1 expansion only exists for expand_ratio != 1
        y = c.conv_bn_act(y, f"{p}.conv.{i}")
        i += 1
    y = c.dw(y, f"{p}.conv.{i}", stride)
    i += 1
    if f"{p}.conv.{i}.norm.weight" not in sd:   # an attention block sits between depthwise and projection
        y = c.attention(y, f"{p}.conv.{i}")
        i += 1
    y = c.conv_bn_act(y, f"{p}.conv.{i}", act=None)
    if stride == 1 and x.shape[1] == y.shape[1]:
        return x + y * residual_scale
    return y


def _dual_path(c, x, p, stride, residual_scale):  # cnn.py:269-380
    sd = c.sd
    r = c.conv_bn_act(x, p + ".residual_path.0")
    r = c.dw(r, p + ".residual_path.1.depthwise", stride)
    r = c.conv_bn_act(r, p + ".residual_path.1.pointwise")
    r = c.conv_bn_act(r, p + ".residual_path.2", act=None)
    d = c.conv_bn_act(x, p + ".dense_path.0")
    d = c.dw(d, p + ".dense_path.1.depthwise", stride)
    d = c.conv_bn_act(d, p + ".dense_path.1.pointwise")
    sc = c.conv_bn_act(x, p + ".shortcut", stride=stride, act=None) if p + ".shortcut.conv.weight" in sd else x
    r = r + sc * residual_scale
    out = c.conv_bn_act(torch.cat([r, d], 1), p + ".fusion")
    if any(k.startswith(p + ".attention.") for k in sd):
        out = c.attention(out, p + ".attention")
    return out


def _wasp(c, x, p, dilations=(1, 6, 12, 18)):  # cnn.py:383-479
    w = torch.softmax(c.sd[p + ".weights"], 0)
    out = c.conv_bn_act(x, p + ".conv1x1") * w[0]
    for i, d in enumerate(dilations):
        out = out + c.conv_bn_act(x, f"{p}.atrous_branches.{i}", dilation=d) * w[i + 1]
    g = c.conv_bn_act(x.mean((2, 3), keepdim=True), p + ".global_branch.1")
    g = F.interpolate(g, size=x.shape[2:], mode="bilinear", align_corners=False)
    out = out + g * w[-1]
    return c.conv_bn_act(out, p + ".fusion")


def heatmaps(kp, hs, sigma):  # common.py:23-51
    coords = torch.arange(hs, dtype=torch.float32, device=kp.device)
    yg, xg = torch.meshgrid(coords, coords, indexing="ij")
    ks = kp * (hs - 1)
    d2 = (xg - ks[..., 0, None, None]) ** 2 + (yg - ks[..., 1, None, None]) ** 2
    hm = torch.exp(-d2 / (2 * sigma ** 2))
    return hm * (kp > 0).all(-1)[..., None, None]


def cnn_forward(sd, cfg, image, depth, kp, train=False, return_stats=False, trace=None, dropout=0.0):
    """cfg: an object / dict with the reference ModelConfig('cnn') attributes.  `trace` (dict) collects the output of
    every ConvBnAct by state-dict prefix (layer-by-layer comparisons in the tests).  `dropout` > 0 (with train=True)
    applies nn.Dropout after every hidden head layer like the reference's training mode (common.py:73-79); parity
    runs keep it at 0."""
    g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    c = _Ctx(sd, train, g("activation"), trace)
    x = torch.cat([image, depth, heatmaps(kp, g("heatmap_size"), g("heatmap_sigma"))], 1)
    x = c.conv_bn_act(x, "conv1.0", stride=g("initial_stride"))
    x = c.conv_bn_act(x, "conv1.1")
    rs = g("residual_scale")
    for i, depth_i in enumerate(g("stage_depths")):
        for j in range(depth_i):
            p = f"stages.{i}.{j}"
            stride = g("stage_strides")[i] if j == 0 else 1
            if p + ".fusion.conv.weight" in sd:
                x = _dual_path(c, x, p, stride, rs)
            else:
                x = _inverted_residual(c, x, p, stride, g("stage_expand_ratios")[i], rs)
    x = _wasp(c, x, "wasp")
    x = F.adaptive_avg_pool2d(x, g("global_pool_size"))
    x = c.conv_bn_act(x, "global_features.1")
    x = c.eca(x, "global_features.2")
    x = x.mean((2, 3))
    n_lin = len(g("regression_dims"))
    for i in range(n_lin):   # dropout is the identity unless asked for: parity runs use eval() / p = 0
        x = _act(F.linear(x, sd[f"pose_head.decoder.{i}.0.weight"], sd[f"pose_head.decoder.{i}.0.bias"]), g("activation"))
        if train and dropout > 0.0:
            x = F.dropout(x, dropout, True)
    x = F.linear(x, sd[f"pose_head.decoder.{n_lin}.weight"], sd[f"pose_head.decoder.{n_lin}.bias"])
    out = x.view(-1, g("num_joints"), 3)
    return (out, c.new_stats) if return_stats else out


def fill_state_dict(sd, seed=0, out_scale_mm=300.0):
    """Deterministic, module-order-independent parameter fill: every tensor is drawn from a generator
    seeded by crc32(key), scaled like the reference's initialisers (kaiming fan_out for conv / linear
    weights, cnn.py:627-639) with BatchNorm affine parameters and running statistics perturbed away
    from (1, 0, 0, 1) so that BN folding is exercised.  Used identically by gen_golden.py (on the live
    reference model) and by the tests (on the product model and this oracle)."""
    out = {}
    for k in sorted(sd):
        v = sd[k]
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) & 0x7FFFFFFF)
        if k.endswith("num_batches_tracked") or k.endswith("x_grid") or k.endswith("y_grid"):
            out[k] = v.clone()
        elif k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif k.endswith("norm.weight") or k.endswith("bn1.weight"):
            out[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.75
        elif k.endswith(".bias"):
            out[k] = torch.randn(v.shape, generator=g) * 0.05
        elif k == "wasp.weights":
            out[k] = torch.randn(v.shape, generator=g) * 0.5
        elif v.dim() >= 2:
            fan_out = v.shape[0] * (v[0][0].numel() if v.dim() > 2 else 1)
            std = math.sqrt(2.0 / fan_out)
            if v.dim() == 3:      # ECA conv1d taps
                std = 0.5
            out[k] = torch.randn(v.shape, generator=g) * std
        else:
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        out[k] = out[k].to(v.dtype)
    last = [k for k in out if k.startswith("pose_head.decoder.") and k.endswith("weight")]
    if last:
        k = sorted(last, key=lambda s: int(s.split(".")[2]))[-1]
        out[k] = out[k] * out_scale_mm      # outputs of pose magnitude (hundreds of mm)
    return out


# ======================================================================================================
# ViT: restated timm VisionTransformer backbone + the reference's fusion / final encoder / head
# ======================================================================================================
def _ln(x, sd, p, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _attend(q, k, v, hd, sdpa, attn_p):
    if sdpa:
        return F.scaled_dot_product_attention(q, k, v, dropout_p=attn_p)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1)
    if attn_p > 0.0:
        a = F.dropout(a, attn_p, True)
    return a @ v


def _mha(sd, p, q_in, kv_in, heads, sdpa=False, attn_p=0.0):
    """nn.MultiheadAttention(batch_first=True) forward: packed in_proj, scaled dot-product attention (dropout on the
    probabilities when attn_p > 0), out_proj.  The averaged attention weights the reference discards are not
    computed.  `sdpa` routes the core through F.scaled_dot_product_attention (fused flash kernels on a GPU)."""
    E = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(q_in, w[:E], b[:E])
    k = F.linear(kv_in, w[E:2 * E], b[E:2 * E])
    v = F.linear(kv_in, w[2 * E:], b[2 * E:])
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    hd = E // heads
    q = q.view(B, Nq, heads, hd).transpose(1, 2)
    k = k.view(B, Nk, heads, hd).transpose(1, 2)
    v = v.view(B, Nk, heads, hd).transpose(1, 2)
    o = _attend(q, k, v, hd, sdpa, attn_p).transpose(1, 2).reshape(B, Nq, E)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def _drop(x, p):
    return F.dropout(x, p, True) if p > 0.0 else x


def _mlp2(sd, p, x, i0, i1, drop=0.0):
    """Linear -> GELU -> Dropout -> Linear -> Dropout (transformers.py:66-72)."""
    h = _drop(F.gelu(F.linear(x, sd[f"{p}.{i0}.weight"], sd[f"{p}.{i0}.bias"])), drop)
    return _drop(F.linear(h, sd[f"{p}.{i1}.weight"], sd[f"{p}.{i1}.bias"]), drop)


def vit_backbone_features(sd, p, x, heads=12, eps=1e-6, sdpa=False):
    """timm 1.0.15 VisionTransformer.forward_features for vit_base_patch16 (published algorithm; timm is not under
    /root/reference and not installed -- SURVEY.md 8c): conv patch embed -> [cls] + tokens + pos_embed -> pre-LN
    blocks (LayerNorm eps 1e-6, fused qkv Linear, 12 heads x 64, exact-erf GELU MLP) -> final LayerNorm."""
    w = sd[p + ".patch_embed.proj.weight"]
    ps = w.shape[-1]
    t = F.conv2d(x, w, sd[p + ".patch_embed.proj.bias"], stride=ps).flatten(2).transpose(1, 2)
    t = torch.cat([sd[p + ".cls_token"].expand(t.shape[0], -1, -1), t], 1) + sd[p + ".pos_embed"]
    E = t.shape[-1]
    hd = E // heads
    i = 0
    while f"{p}.blocks.{i}.norm1.weight" in sd:
        bp = f"{p}.blocks.{i}"
        h = _ln(t, sd, bp + ".norm1", eps)
        qkv = F.linear(h, sd[bp + ".attn.qkv.weight"], sd[bp + ".attn.qkv.bias"])
        B, N, _ = qkv.shape
        q, k, v = qkv.view(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
        o = _attend(q, k, v, hd, sdpa, 0.0).transpose(1, 2).reshape(B, N, E)
        t = t + F.linear(o, sd[bp + ".attn.proj.weight"], sd[bp + ".attn.proj.bias"])
        h = _ln(t, sd, bp + ".norm2", eps)
        t = t + F.linear(F.gelu(F.linear(h, sd[bp + ".mlp.fc1.weight"], sd[bp + ".mlp.fc1.bias"])),
                         sd[bp + ".mlp.fc2.weight"], sd[bp + ".mlp.fc2.bias"])
        i += 1
    return _ln(t, sd, p + ".norm", eps)


def vit_forward(sd, cfg, image, depth, kp, train=False, sdpa=False):
    """TransformerPoseEstimation.forward (reference: src/models/transformers.py:326-373).  Default: eval mode (every
    dropout off).  train=True applies the reference's dropout sites with the config's rates (attention probabilities,
    after each attention / MLP Linear, head) -- the timm backbone has none; `sdpa` uses the fused attention kernels."""
    g = (lambda k: cfg[k]) if isinstance(cfg, dict) else (lambda k: getattr(cfg, k))
    heads = g("transformer_heads")
    dp = float(g("transformer_dropout_rate")) if train else 0.0
    ap = float(g("transformer_attention_dropout_rate")) if train else 0.0
    hp = float(g("regression_dropout")) if train else 0.0
    tok = vit_backbone_features(sd, "vit_backbone", torch.cat([image, depth], 1), sdpa=sdpa)[:, 1:]   # drop the cls prefix token
    hm = heatmaps(kp, g("heatmap_size"), g("heatmap_sigma"))
    w = sd["heatmap_patch_embed.proj.weight"]
    hm_tok = F.conv2d(hm, w, sd["heatmap_patch_embed.proj.bias"], stride=w.shape[-1]).flatten(2).transpose(1, 2)
    hm_tok = hm_tok + sd["pos_embed_hm"]
    x_img, x_hm = tok, hm_tok
    for i in range(g("num_cross_modal_layers")):     # CrossModalFusionBlock, transformers.py:85-137 (LayerNorm eps 1e-5)
        p = f"cross_modal_fusion_layers.{i}"
        x_img = x_img + _drop(_mha(sd, p + ".cross_attn_img_to_hm", _ln(x_img, sd, p + ".norm_img_q", 1e-5),
                                   _ln(x_hm, sd, p + ".norm_hm_kv", 1e-5), heads, sdpa, ap), dp)
        x_hm = x_hm + _drop(_mha(sd, p + ".cross_attn_hm_to_img", _ln(x_hm, sd, p + ".norm_hm_q", 1e-5),
                                 _ln(x_img, sd, p + ".norm_img_kv", 1e-5), heads, sdpa, ap), dp)
        x_img = x_img + _mlp2(sd, p + ".mlp_img", _ln(x_img, sd, p + ".norm_img_mlp", 1e-5), 0, 3, dp)
        x_hm = x_hm + _mlp2(sd, p + ".mlp_hm", _ln(x_hm, sd, p + ".norm_hm_mlp", 1e-5), 0, 3, dp)
    t = torch.cat([sd["final_cls_token"].expand(x_img.shape[0], -1, -1), x_img, x_hm], 1) + sd["final_pos_embed"]
    for i in range(g("final_encoder_depth")):        # TransformerEncoderBlock, transformers.py:49-82
        p = f"final_encoder.{i}"
        h = _ln(t, sd, p + ".norm1", 1e-5)
        t = t + _drop(_mha(sd, p + ".attn", h, h, heads, sdpa, ap), dp)
        t = t + _mlp2(sd, p + ".mlp", _ln(t, sd, p + ".norm2", 1e-5), 0, 3, dp)
    x = _ln(t[:, 0], sd, "norm_out", 1e-5)
    dims = g("regression_hidden_dims")
    for i in range(len(dims)):                       # transformers.py:20-26: Linear at decoder.{0,3,6,..}
        x = _drop(F.gelu(F.linear(x, sd[f"pose_head.decoder.{3 * i}.weight"], sd[f"pose_head.decoder.{3 * i}.bias"])), hp)
    n = 3 * len(dims)
    x = F.linear(x, sd[f"pose_head.decoder.{n}.weight"], sd[f"pose_head.decoder.{n}.bias"])
    return x.view(-1, g("num_joints"), 3)


def composite_loss(pred, gt, w_mse=1.0, w_l1=1.0, w_ij=100.0, w_root=1.0):
    """ComprehensivePoseLoss.forward (reference: src/loss.py:57-85, weights config.py:15-18) in plain torch ops."""
    d = pred - gt
    J = pred.shape[1]
    iu = torch.triu_indices(J, J, 1, device=pred.device)

    def pd(t):
        return torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]
    return (w_mse * (d ** 2).mean() + w_l1 * d.abs().mean() + w_ij * (pd(pred) - pd(gt)).abs().mean()
            + w_root * d[:, 0].abs().mean())


def fill_vit_state_dict(sd, seed=0, out_scale_mm=300.0):
    """Deterministic key-seeded fill for the ViT (same idea as fill_state_dict): trunc-normal-like 0.02 embeddings,
    xavier-scaled Linear weights, LayerNorm affine parameters perturbed away from (1, 0)."""
    out = {}
    for k in sorted(sd):
        v = sd[k]
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + seed) & 0x7FFFFFFF)
        if k.endswith("x_grid") or k.endswith("y_grid"):
            out[k] = v.clone()
        elif "norm" in (k.split(".") + [""])[-2] and k.endswith("weight") and v.dim() == 1:
            out[k] = torch.rand(v.shape, generator=g) * 0.4 + 0.8
        elif v.dim() == 1:
            out[k] = torch.randn(v.shape, generator=g) * 0.02
        elif "pos_embed" in k or "cls_token" in k:
            out[k] = torch.randn(v.shape, generator=g) * 0.02
        else:
            fan_in = v[0].numel()
            fan_out = v.shape[0]
            out[k] = torch.randn(v.shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))
        out[k] = out[k].to(v.dtype)
    last = sorted([k for k in out if k.startswith("pose_head.decoder.") and k.endswith("weight")],
                  key=lambda s: int(s.split(".")[2]))[-1]
    out[last] = out[last] * out_scale_mm * 10.0
    return out
