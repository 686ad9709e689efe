"""TransformerPoseEstimation on B200 (reference: src/models/transformers.py:7-373).

The module tree (attribute names, parameter shapes) reproduces the reference's -- including the timm
``VisionTransformer`` it wraps as ``vit_backbone`` (timm 1.0.15 is a third-party dependency that is not under
/root/reference; its ViT-B/16 forward is restated from the published algorithm) -- so
``state_dict()`` keys match and reference checkpoints load unchanged.  The modules are parameter containers;
``forward`` runs hand-written sm_100a kernels through the C ABI: every Linear / patch-embedding on the
tcgen05 GEMM (bias, GELU and the residual add fused in the epilogue), LayerNorm / attention / token assembly
as bandwidth kernels, bf16 activations with fp32 statistics.  In training mode the same plan keeps the
activations the backward pass needs and ``backward`` walks the model in reverse: data gradients and weight
gradients are tcgen05 GEMMs that read the saved activations and the weights in place (MN-major operands),
weight gradients accumulate into the flat fp32 ``.grad`` buffer (params.FlatParams).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _lib
from ..params import FlatParams
from ..utils import GraphedForward, wgrad_splits, activation_id, get_activation
from .common import GaussianHeatmapGenerator

_VIT_TABLE = {  # timm model name -> (embed_dim, depth, heads, patch)
    "vit_base_patch16_384": (768, 12, 12, 16),
    "vit_base_patch16_224": (768, 12, 12, 16),
    "vit_small_patch16_224": (384, 12, 6, 16),
    "vit_large_patch16_224": (1024, 24, 16, 16),
}


# ------------------------------------------------------------------------------------------------------
# parameter containers
# ------------------------------------------------------------------------------------------------------
class PoseRegressionHead(nn.Module):  # transformers.py:7-31 (flat Sequential: decoder.{0,3,6,..})
    def __init__(self, in_features, num_joints, hidden_dims=(512, 256), dropout=0.2, activation="gelu"):
        super().__init__()
        self.num_joints = num_joints
        self.activation = activation
        layers = []
        prev = in_features
        for h in hidden_dims:
            layers += [nn.Linear(prev, h), get_activation(activation), nn.Dropout(dropout)]
            prev = h
        layers.append(nn.Linear(prev, num_joints * 3))
        self.decoder = nn.Sequential(*layers)

    def linears(self):
        return [m for m in self.decoder if isinstance(m, nn.Linear)]

    def forward(self, x):  # transformers.py:28-31 (standalone use; the model's plan runs the same GEMMs in its own buffers)
        from ..ops import mlp_head_forward
        drops = [m.p for m in self.decoder if isinstance(m, nn.Dropout)]
        pose = mlp_head_forward(x.reshape(x.size(0), -1), self.linears(), self.activation,
                                drops[0] if (self.training and drops) else 0.0)
        return pose.view(-1, self.num_joints, 3)


class PatchEmbedding(nn.Module):  # transformers.py:33-47
    def __init__(self, img_size_h, img_size_w, patch_size, in_chans, embed_dim):
        super().__init__()
        if img_size_h % patch_size != 0 or img_size_w % patch_size != 0:
            raise ValueError(f"Image dims ({img_size_h}x{img_size_w}) must be divisible by patch size ({patch_size}).")
        self.num_patches = (img_size_h // patch_size) * (img_size_w // patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def _mlp(embed_dim, hidden, dropout, activation):
    return nn.Sequential(nn.Linear(embed_dim, hidden), get_activation(activation), nn.Dropout(dropout),
                         nn.Linear(hidden, embed_dim), nn.Dropout(dropout))


class TransformerEncoderBlock(nn.Module):  # transformers.py:49-82
    def __init__(self, embed_dim, num_heads, mlp_ratio, dropout_rate, attention_dropout_rate, activation="gelu"):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim)
        self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=attention_dropout_rate, batch_first=True)
        self.attn_dropout = nn.Dropout(dropout_rate)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.mlp = _mlp(embed_dim, int(embed_dim * mlp_ratio), dropout_rate, activation)


class CrossModalFusionBlock(nn.Module):  # transformers.py:85-137
    def __init__(self, embed_dim, num_heads, mlp_ratio, dropout_rate, attention_dropout_rate, activation="gelu"):
        super().__init__()
        self.norm_img_q = nn.LayerNorm(embed_dim)
        self.norm_hm_kv = nn.LayerNorm(embed_dim)
        self.cross_attn_img_to_hm = nn.MultiheadAttention(embed_dim, num_heads, dropout=attention_dropout_rate,
                                                          batch_first=True)
        self.dropout_img = nn.Dropout(dropout_rate)
        self.norm_hm_q = nn.LayerNorm(embed_dim)
        self.norm_img_kv = nn.LayerNorm(embed_dim)
        self.cross_attn_hm_to_img = nn.MultiheadAttention(embed_dim, num_heads, dropout=attention_dropout_rate,
                                                          batch_first=True)
        self.dropout_hm = nn.Dropout(dropout_rate)
        hidden = int(embed_dim * mlp_ratio)
        self.norm_img_mlp = nn.LayerNorm(embed_dim)
        self.mlp_img = _mlp(embed_dim, hidden, dropout_rate, activation)
        self.norm_hm_mlp = nn.LayerNorm(embed_dim)
        self.mlp_hm = _mlp(embed_dim, hidden, dropout_rate, activation)


class _VitAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)


class _VitMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


class _VitBlock(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _VitAttention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _VitMlp(dim, dim * 4)


class _VitPatchEmbed(nn.Module):
    def __init__(self, img_size, patch, in_chans, dim):
        super().__init__()
        self.num_patches = (img_size[0] // patch) * (img_size[1] // patch)
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)


class VisionTransformerBackbone(nn.Module):
    """Container with timm ``VisionTransformer``'s state-dict layout (cls_token, pos_embed, patch_embed.proj,
    blocks.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}, norm) and its non-pretrained initialisation
    (trunc-normal 0.02 weights, zero biases, cls_token normal 1e-6)."""

    def __init__(self, name, img_size, in_chans):
        super().__init__()
        if name not in _VIT_TABLE:
            raise NotImplementedError(f"vit_model_name {name!r}: known {sorted(_VIT_TABLE)}")
        dim, depth, heads, patch = _VIT_TABLE[name]
        if img_size[0] % patch or img_size[1] % patch:
            raise ValueError(f"image_size {img_size} must be divisible by the patch size {patch}")
        self.embed_dim, self.num_heads, self.patch_size = dim, heads, patch
        self.num_prefix_tokens = 1
        self.patch_embed = _VitPatchEmbed(img_size, patch, in_chans, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, dim))
        self.blocks = nn.Sequential(*[_VitBlock(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


def adapt_timm_vit_state_dict(sd, in_chans, num_patches_hw, num_prefix_tokens=1):
    """A timm ``VisionTransformer`` state dict (3-channel patch embedding, the position embedding of its pre-training
    resolution) -> the tensors of ``VisionTransformerBackbone`` for this model:
      * patch_embed.proj.weight 3 -> `in_chans` channels exactly like the reference (src/models/transformers.py:179-214:
        the RGB filters are kept, every extra channel gets their mean; fewer channels: the mean repeated);
      * pos_embed resampled to the new patch grid the way timm does when ``img_size`` differs (``resample_abs_pos_embed``:
        prefix tokens kept, the grid interpolated bicubically with antialiasing);
      * classifier head / pre-logits entries dropped (``num_classes=0``)."""
    out = {k: v for k, v in sd.items() if not k.startswith(("head.", "fc_norm.", "head_drop."))}
    w = out["patch_embed.proj.weight"]
    c0 = w.shape[1]
    if in_chans > c0:
        extra = w.mean(dim=1, keepdim=True).repeat(1, in_chans - c0, 1, 1)
        out["patch_embed.proj.weight"] = torch.cat([w, extra], dim=1)
    elif in_chans < c0:
        out["patch_embed.proj.weight"] = w.mean(dim=1, keepdim=True).repeat(1, in_chans, 1, 1)
    pe = out["pos_embed"]
    gh, gw = num_patches_hw
    n_old = pe.shape[1] - num_prefix_tokens
    if n_old != gh * gw:
        g0 = int(round(n_old ** 0.5))
        if g0 * g0 != n_old:
            raise ValueError(f"pos_embed with {n_old} grid tokens is not a square grid")
        prefix, grid = pe[:, :num_prefix_tokens], pe[:, num_prefix_tokens:]
        grid = grid.reshape(1, g0, g0, -1).permute(0, 3, 1, 2).float()
        grid = torch.nn.functional.interpolate(grid, size=(gh, gw), mode="bicubic", antialias=True)
        grid = grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1).to(pe.dtype)
        out["pos_embed"] = torch.cat([prefix, grid], dim=1)
    return out


def _pretrained_vit_state_dict(name):
    """vit_pretrained=True (the reference's default, transformers.py:174-179): the weights come from a local file named by
    POSE_VIT_WEIGHTS (a ``torch.save``d timm state dict -- there is no network on the training boxes) or, when timm is
    installed, from ``timm.create_model(name, pretrained=True)``."""
    import os
    path = os.environ.get("POSE_VIT_WEIGHTS")
    if path:
        sd = torch.load(path, map_location="cpu", weights_only=True)
        return sd.get("state_dict", sd.get("model", sd)) if isinstance(sd, dict) else sd
    try:
        import timm
    except ImportError:
        raise NotImplementedError(
            "vit_pretrained=True needs the timm weights: point POSE_VIT_WEIGHTS at a saved timm state dict of "
            f"{name!r} (or install timm), or construct with vit_pretrained=False and load a checkpoint") from None
    return timm.create_model(name, pretrained=True, num_classes=0).state_dict()


class TransformerPoseEstimation(nn.Module):  # transformers.py:140-373
    def __init__(self, config):
        super().__init__()
        c = config
        self.config = c
        self.vit_backbone = VisionTransformerBackbone(c.vit_model_name, tuple(c.image_size), c.image_in_channels)
        if getattr(c, "vit_pretrained", False):
            p = self.vit_backbone.patch_size
            sd = adapt_timm_vit_state_dict(_pretrained_vit_state_dict(c.vit_model_name), c.image_in_channels,
                                           (c.image_size[0] // p, c.image_size[1] // p))
            self.vit_backbone.load_state_dict(sd)
        if getattr(c, "vit_freeze_backbone", False):
            # transformers.py:226-236: everything in the backbone but the patch embedding that was adapted to a different
            # channel count (a pre-trained 3-channel embedding widened to RGB-D stays trainable)
            adapted = c.image_in_channels != 3
            for pname, prm in self.vit_backbone.named_parameters():
                if not (pname.startswith("patch_embed.proj") and adapted):
                    prm.requires_grad = False
        c.transformer_embed_dim = self.vit_backbone.embed_dim
        E = c.transformer_embed_dim
        self.heatmap_generator = GaussianHeatmapGenerator(c.num_joints, c.heatmap_size, c.heatmap_sigma)
        self.heatmap_patch_embed = PatchEmbedding(c.heatmap_size, c.heatmap_size, c.heatmap_patch_size,
                                                  c.heatmap_in_channels, E)
        self.pos_embed_hm = nn.Parameter(torch.zeros(1, self.heatmap_patch_embed.num_patches, E))
        blk = (E, c.transformer_heads, c.transformer_mlp_ratio, c.transformer_dropout_rate,
               c.transformer_attention_dropout_rate, c.activation)
        self.cross_modal_fusion_layers = nn.ModuleList([CrossModalFusionBlock(*blk)
                                                        for _ in range(c.num_cross_modal_layers)])
        self.final_cls_token = nn.Parameter(torch.zeros(1, 1, E))
        n_final = 1 + self.vit_backbone.patch_embed.num_patches + self.heatmap_patch_embed.num_patches
        self.final_pos_embed = nn.Parameter(torch.zeros(1, n_final, E))
        self.pos_drop = nn.Dropout(c.transformer_dropout_rate)
        self.final_pos_drop = nn.Dropout(c.transformer_dropout_rate)
        self.final_encoder = nn.ModuleList([TransformerEncoderBlock(*blk) for _ in range(c.final_encoder_depth)])
        self.norm_out = nn.LayerNorm(E)
        self.pose_head = PoseRegressionHead(E, c.num_joints, c.regression_hidden_dims, c.regression_dropout,
                                            c.activation)
        self._initialize_weights()
        self._plans = {}

    def _initialize_weights(self):  # transformers.py:307-324
        nn.init.trunc_normal_(self.pos_embed_hm, std=0.02)
        nn.init.trunc_normal_(self.final_pos_embed, std=0.02)
        nn.init.trunc_normal_(self.final_cls_token, std=0.02)
        for mod in (self.heatmap_patch_embed, self.cross_modal_fusion_layers, self.final_encoder, self.pose_head,
                    self.norm_out):
            mod.apply(self._init_weights_for_linear)

    @staticmethod
    def _init_weights_for_linear(m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def plan(self, B, device):
        key = (B, device.index)
        p = self._plans.get(key)
        if p is None or not p.flat.intact():
            p = VitPlan(self, B, device)
            self._plans[key] = p
        return p

    def forward(self, image, depth, keypoints_2d):
        """image [B,3,H,W], depth [B,1,H,W], keypoints_2d [B,J,2] (fp32, CUDA) -> joints [B,J,3] fp32."""
        plan = self.plan(image.shape[0], image.device)
        if self.training and torch.is_grad_enabled():
            return _VitTrainFn.apply(plan, image, depth, keypoints_2d, self.final_cls_token)
        if not self.training and image.shape[0] <= GraphedForward.MAX_BATCH:
            # eval forward of a small batch: launch bound -> replayed from one CUDA graph (the bf16 shadow weights are
            # refreshed outside the graph when a parameter changed in place)
            if plan.graphed is None:
                plan.graphed = GraphedForward(lambda i, d, k: plan.forward(i, d, k, save=False, training=False),
                                              prepare=plan.flat.refresh_shadow)
            out = plan.graphed(image, depth, keypoints_2d)
        else:
            out = plan.forward(image, depth, keypoints_2d, save=False, training=self.training)
        return out.view(-1, self.config.num_joints, 3).clone()


class _VitTrainFn(torch.autograd.Function):
    """Autograd node of the whole model: backward runs the plan's reverse pass, which accumulates straight into
    the parameters' flat ``.grad`` views (the `anchor` parameter only ties the node into the graph)."""

    @staticmethod
    def forward(ctx, plan, image, depth, kp, anchor):
        ctx.plan = plan
        out = plan.forward(image, depth, kp, save=True)
        return out.view(-1, plan.J, 3).clone()

    @staticmethod
    def backward(ctx, dout):
        ctx.plan.backward(dout.contiguous())
        return None, None, None, None, None


# ------------------------------------------------------------------------------------------------------
# launch plan: forward (eval or training) and backward at a fixed batch size
# ------------------------------------------------------------------------------------------------------
class VitPlan:
    def __init__(self, model: TransformerPoseEstimation, B: int, device):
        self.model, self.B, self.dev = model, B, device
        self.lib = _lib.lib()
        self._sp = None          # stream handle of the current forward / backward (see call)
        self.graphed = None      # GraphedForward of the eval forward (small batches)
        c = model.config
        self.J = c.num_joints
        self.E = c.transformer_embed_dim
        self.H, self.W = int(c.image_size[0]), int(c.image_size[1])
        bb = model.vit_backbone
        self.P = bb.patch_size
        self.T_img = bb.patch_embed.num_patches
        self.T_hm = model.heatmap_patch_embed.num_patches
        self.T_fin = 1 + self.T_img + self.T_hm
        self.act = activation_id(c.activation)
        if self.act not in (2, 3):
            raise NotImplementedError("transformer activation: gelu or silu")
        # nn.Dropout sites of the reference (transformers.py:24,61-72,98-124,363): active in training mode only; the timm
        # backbone has no dropout (all its drop rates default to 0)
        self.p_drop = float(c.transformer_dropout_rate)
        self.p_attn = float(c.transformer_attention_dropout_rate)
        self.p_head = float(c.regression_dropout)
        self.step_count = 0
        self.seeds = {}
        self.training = False
        self.flat = FlatParams.of(model.parameters())
        self.bufs = {}
        self.launches = 0
        self.keep = []

    # ---- small helpers -------------------------------------------------------------------------------
    def buf(self, name, *shape, dtype=torch.bfloat16, zero=False):
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=self.dev)
            self.bufs[name] = t
        elif zero:
            t.zero_()
        return t

    def call(self, name, *args):
        # the stream handle is looked up once per forward / backward (a property chain in torch, ~3 us: it was the largest
        # single item of the per-launch host cost of a step with 500-600 launches)
        rc = getattr(self.lib, name)(*args, self._sp if self._sp is not None else _lib.stream_ptr())
        if rc:
            _lib.check(rc, name)
        self.launches += 1

    def seed(self, site):
        """Dropout seed of a site for the current step (the backward pass regenerates the mask from it)."""
        idx = self.seeds.get(site)
        if idx is None:
            idx = self.seeds[site] = len(self.seeds) + 1
        # with a bound pose_step_state the per-step part of the seed lives in device memory (graph replay)
        return idx if _lib.step_state_bound() else self.step_count * 4096 + idx

    def drop(self, site, p):
        """(p, seed) for a dropout site, or None when dropout is off (eval mode / p = 0)."""
        return (p, self.seed(site)) if (self.training and p > 0.0) else None

    def _epi(self, out, ldc, bias=None, act=0, residual=None, ldr=0, preact=None, accumulate=0, out_scale=1.0,
             res_scale=1.0, drop=None):
        e = _lib.PoseGemmEpilogue()
        e.bias = bias.data_ptr() if bias is not None else None
        e.residual = residual.data_ptr() if residual is not None else None
        e.C = out.data_ptr()
        e.ldc, e.ldr, e.act = ldc, ldr, act
        e.out_dtype = 0 if out.dtype == torch.float32 else 1
        e.out_scale, e.res_scale = out_scale, res_scale
        e.preact = preact.data_ptr() if preact is not None else None
        e.accumulate = accumulate
        if drop is not None:
            e.drop_p, e.drop_seed = drop
        return e

    # ---- forward ops ---------------------------------------------------------------------------------
    def linear(self, name, x, M, weight, bias, rows=None, act=0, residual=None, preact=False, fp32=False, drop=None):
        """y = act(x @ W[rows]^T + b[rows]) (+ residual); returns y (and keeps the pre-activation when asked)."""
        w16 = self.flat.w16(weight, rows)
        N, K = w16.shape
        b = self.flat.f32(bias)
        if rows is not None:
            b = b[rows[0]:rows[1]]
        y = self.buf(name, M, N, dtype=torch.float32 if fp32 else torch.bfloat16)
        u = self.buf(name + ".u", M, N) if preact else None
        e = self._epi(y, N, b, act, residual, N if residual is not None else 0, u, drop=drop)
        self.call("pose_gemm_bf16_ex", x.data_ptr(), K, w16.data_ptr(), K, M, N, K, C.byref(e))
        return (y, u) if preact else y

    def layernorm(self, name, x, ln, M, rows=None, in_group=0, in_off=0):
        """LayerNorm of M rows; with `rows`, row r = g*rows+i reads x row g*in_group+in_off+i (token sub-ranges)."""
        D = self.E
        y = self.buf(name, M, D)
        if rows is None:
            rows, in_group = M, M
        self.call("pose_layernorm_bf16", x.data_ptr(), self.flat.f32(ln.weight).data_ptr(),
                  self.flat.f32(ln.bias).data_ptr(), float(ln.eps), M, rows, in_group, in_off, rows, 0, D, y.data_ptr())
        return y

    def attention(self, name, q, k, v, Nq, Nk, heads, ldq, ldk, ldv, save, drop=None):
        """softmax(q k^T / sqrt(hd)) v per (sample, head); q/k/v are column offsets into row-major buffers."""
        E, B = self.E, self.B
        hd = E // heads
        o = self.buf(name, B * Nq, E)
        lse = self.buf(name + ".lse", B, heads, Nq, dtype=torch.float32) if save else None
        self.call("pose_attention_bf16", q, k, v, o.data_ptr(), B, heads, Nq, Nk, hd, ldq, ldk, ldv, E, Nq * ldq,
                  Nk * ldk, Nk * ldv, Nq * E, 1.0 / math.sqrt(hd), lse.data_ptr() if save else None,
                  drop[0] if drop else 0.0, drop[1] if drop else 0)
        return o

    def encoder_block(self, pre, x, T, norm1, wqkv, bqkv, wo, bo, norm2, fc1, fc2, heads, save, p_drop=0.0, p_attn=0.0,
                      act=None):
        """pre-LN transformer block (timm Block and TransformerEncoderBlock, transformers.py:75-82); dropout on the
        attention weights, on the projected attention output, after the activation and after fc2 (transformers.py:61-72)."""
        M, E = self.B * T, self.E
        h = self.layernorm(pre + "h", x, norm1, M)
        qkv = self.linear(pre + "qkv", h, M, wqkv, bqkv)
        p = qkv.data_ptr()
        o = self.attention(pre + "o", p, p + 2 * E, p + 4 * E, T, T, heads, 3 * E, 3 * E, 3 * E, save,
                           drop=self.drop(pre + "attn", p_attn))
        x1 = self.linear(pre + "x1", o, M, wo, bo, residual=x, drop=self.drop(pre + "proj", p_drop))
        h2 = self.layernorm(pre + "h2", x1, norm2, M)
        g = self.linear(pre + "g", h2, M, fc1.weight, fc1.bias, act=self.act if act is None else act, preact=save,
                        drop=self.drop(pre + "act", p_drop))
        if save:
            g = g[0]
        return self.linear(pre + "x2", g, M, fc2.weight, fc2.bias, residual=x1, drop=self.drop(pre + "fc2", p_drop))

    def mlp_residual(self, pre, x, M, norm, mlp, save):
        h = self.layernorm(pre + "h", x, norm, M)
        g = self.linear(pre + "g", h, M, mlp[0].weight, mlp[0].bias, act=self.act, preact=save,
                        drop=self.drop(pre + "act", self.p_drop))
        if save:
            g = g[0]
        return self.linear(pre + "y", g, M, mlp[3].weight, mlp[3].bias, residual=x, drop=self.drop(pre + "fc2", self.p_drop))

    def cross_block(self, pre, blk, x_img, x_hm, save):
        B, E, Ti, Th = self.B, self.E, self.T_img, self.T_hm
        heads = self.model.config.transformer_heads
        Mi, Mh = B * Ti, B * Th
        a1, a2 = blk.cross_attn_img_to_hm, blk.cross_attn_hm_to_img
        img_q = self.layernorm(pre + "img_q", x_img, blk.norm_img_q, Mi)
        hm_kv = self.layernorm(pre + "hm_kv", x_hm, blk.norm_hm_kv, Mh)
        q1 = self.linear(pre + "q1", img_q, Mi, a1.in_proj_weight, a1.in_proj_bias, rows=(0, E))
        kv1 = self.linear(pre + "kv1", hm_kv, Mh, a1.in_proj_weight, a1.in_proj_bias, rows=(E, 3 * E))
        o1 = self.attention(pre + "o1", q1.data_ptr(), kv1.data_ptr(), kv1.data_ptr() + 2 * E, Ti, Th, heads, E, 2 * E,
                            2 * E, save, drop=self.drop(pre + "attn1", self.p_attn))
        x_img1 = self.linear(pre + "x_img1", o1, Mi, a1.out_proj.weight, a1.out_proj.bias, residual=x_img,
                             drop=self.drop(pre + "proj1", self.p_drop))
        hm_q = self.layernorm(pre + "hm_q", x_hm, blk.norm_hm_q, Mh)
        img_kv = self.layernorm(pre + "img_kv", x_img1, blk.norm_img_kv, Mi)
        q2 = self.linear(pre + "q2", hm_q, Mh, a2.in_proj_weight, a2.in_proj_bias, rows=(0, E))
        kv2 = self.linear(pre + "kv2", img_kv, Mi, a2.in_proj_weight, a2.in_proj_bias, rows=(E, 3 * E))
        o2 = self.attention(pre + "o2", q2.data_ptr(), kv2.data_ptr(), kv2.data_ptr() + 2 * E, Th, Ti, heads, E, 2 * E,
                            2 * E, save, drop=self.drop(pre + "attn2", self.p_attn))
        x_hm1 = self.linear(pre + "x_hm1", o2, Mh, a2.out_proj.weight, a2.out_proj.bias, residual=x_hm,
                            drop=self.drop(pre + "proj2", self.p_drop))
        x_img2 = self.mlp_residual(pre + "mi.", x_img1, Mi, blk.norm_img_mlp, blk.mlp_img, save)
        x_hm2 = self.mlp_residual(pre + "mh.", x_hm1, Mh, blk.norm_hm_mlp, blk.mlp_hm, save)
        return x_img2, x_hm2

    def forward(self, image, depth, kp, save, training=None):
        self._sp = _lib.stream_ptr()
        m, B, E = self.model, self.B, self.E
        c = m.config
        _lib.require_cuda(image, "image", torch.float32)
        _lib.require_cuda(depth, "depth", torch.float32)
        _lib.require_cuda(kp, "keypoints_2d", torch.float32)
        if tuple(image.shape) != (B, 3, self.H, self.W) or tuple(depth.shape) != (B, 1, self.H, self.W) or \
                tuple(kp.shape) != (B, self.J, 2) or c.image_in_channels != 4:
            raise ValueError(f"expected image [{B},3,{self.H},{self.W}], depth [{B},1,{self.H},{self.W}], "
                             f"keypoints [{B},{self.J},2]")
        # dropout follows the module's mode (the reference applies it in train mode with or without autograd);
        # `save` only decides whether the backward's operands are kept
        self.training = bool(save) if training is None else bool(training)
        if self.training:
            self.step_count += 1
        self.flat.refresh_shadow()
        self.launches = 0
        self.saved = save
        bb = m.vit_backbone
        Ti, Th, P = self.T_img, self.T_hm, self.P
        # ---- image / depth stream: timm VisionTransformer.forward_features -------------------------
        pimg = self.buf("pimg", B * Ti, 4 * P * P)
        self.call("pose_patchify_bf16", image.data_ptr(), 3, depth.data_ptr(), 1, B, self.H, self.W, P, pimg.data_ptr())
        tok = self.linear("tok", pimg, B * Ti, bb.patch_embed.proj.weight, bb.patch_embed.proj.bias)
        x = self.buf("bb.x0", B * (Ti + 1), E)
        self.call("pose_token_concat_bf16", x.data_ptr(), B, Ti + 1, E, self.flat.f32(bb.cls_token).data_ptr(),
                  tok.data_ptr(), Ti, None, 0, self.flat.f32(bb.pos_embed).data_ptr())
        for i, blk in enumerate(bb.blocks):
            x = self.encoder_block(f"bb{i}.", x, Ti + 1, blk.norm1, blk.attn.qkv.weight, blk.attn.qkv.bias,
                                   blk.attn.proj.weight, blk.attn.proj.bias, blk.norm2, blk.mlp.fc1, blk.mlp.fc2,
                                   bb.num_heads, save, act=3)   # timm's Mlp is exact GELU whatever config.activation says
        self.bb_out = x
        # final norm of the patch tokens only: the prefix (cls) token is dropped (transformers.py:336-346)
        x_img = self.layernorm("x_img", x, bb.norm, B * Ti, rows=Ti, in_group=Ti + 1, in_off=1)
        # ---- heat-map stream ---------------------------------------------------------------------------
        hs, hp = int(c.heatmap_size), int(c.heatmap_patch_size)
        phm = self.buf("phm", B * Th, self.J * hp * hp)
        if hp % 8 == 0:
            # the Gaussians are rendered straight into the patch-embedding GEMM's A operand (no fp32 planes in HBM)
            self.call("pose_heatmap_patchify_bf16", kp.data_ptr(), B, self.J, hs, float(c.heatmap_sigma), hp, phm.data_ptr())
        else:
            hm = self.buf("hm", B, self.J, hs, hs, dtype=torch.float32)
            self.call("pose_heatmap_render", kp.data_ptr(), B, self.J, hs, float(c.heatmap_sigma), hm.data_ptr(), 0, 0, 0, 0)
            self.call("pose_patchify_bf16", hm.data_ptr(), self.J, None, 0, B, hs, hs, hp, phm.data_ptr())
        hpe = m.heatmap_patch_embed.proj
        hm_tok = self.linear("hm_tok", phm, B * Th, hpe.weight, hpe.bias)
        x_hm = self.buf("x_hm", B * Th, E)
        self.call("pose_token_concat_bf16", x_hm.data_ptr(), B, Th, E, None, hm_tok.data_ptr(), Th, None, 0,
                  self.flat.f32(m.pos_embed_hm).data_ptr())
        # ---- cross-modal fusion ------------------------------------------------------------------------
        for i, blk in enumerate(m.cross_modal_fusion_layers):
            x_img, x_hm = self.cross_block(f"cm{i}.", blk, x_img, x_hm, save)
        # ---- final encoder -----------------------------------------------------------------------------
        Tf = self.T_fin
        t = self.buf("fin.x0", B * Tf, E)
        self.call("pose_token_concat_bf16", t.data_ptr(), B, Tf, E, self.flat.f32(m.final_cls_token).data_ptr(),
                  x_img.data_ptr(), Ti, x_hm.data_ptr(), Th, self.flat.f32(m.final_pos_embed).data_ptr())
        d = self.drop("final_pos_drop", self.p_drop)              # final_pos_drop (transformers.py:363)
        if d is not None:
            self.call("pose_dropout_bf16", t.data_ptr(), t.numel(), d[0], d[1], t.data_ptr())
        for i, blk in enumerate(m.final_encoder):
            t = self.encoder_block(f"fe{i}.", t, Tf, blk.norm1, blk.attn.in_proj_weight, blk.attn.in_proj_bias,
                                   blk.attn.out_proj.weight, blk.attn.out_proj.bias, blk.norm2, blk.mlp[0], blk.mlp[3],
                                   c.transformer_heads, save, self.p_drop, self.p_attn)
        self.fin_out = t
        h = self.layernorm("cls_out", t, m.norm_out, B, rows=1, in_group=Tf, in_off=0)
        lins = m.pose_head.linears()
        for i, lin in enumerate(lins):
            last = i == len(lins) - 1
            h = self.linear(f"head{i}", h, B, lin.weight, lin.bias, act=0 if last else self.act,
                            preact=save and not last, fp32=last, drop=None if last else self.drop(f"head{i}", self.p_head))
            if save and not last:
                h = h[0]
        return h

    # ---- backward ops --------------------------------------------------------------------------------
    def _splits(self, n_out, n_in, m_rows):
        return wgrad_splits(n_out, n_in, m_rows)

    def drop_grad(self, name, d, site, p):
        """Gradient entering a dropout site: the forward's mask (same seed) applied to d -> a new buffer."""
        dd = self.drop(site, p)
        if dd is None:
            return d
        out = self.buf(name, *d.shape)
        self.call("pose_dropout_bf16", d.data_ptr(), d.numel(), dd[0], dd[1], out.data_ptr())
        return out

    def linear_bwd(self, dy, ldy, M, x, weight, bias, rows=None, dx_name=None, act_u=None, act_drop=None):
        """Backward of y = x @ W[rows]^T + b[rows] given dy [M, N] (pitch ldy): accumulates dW and db, returns
        dx = dy @ W (times the saved derivative act'(u) `act_u`, and the dropout mask `act_drop`, when the layer's INPUT x
        was drop(act(u))) or None."""
        w16 = self.flat.w16(weight, rows)
        N, K = w16.shape
        if weight.requires_grad:                 # frozen layers (vit_freeze_backbone) only pass the data gradient on
            gw = self.flat.g32(weight, rows).view(N, K)
            e = self._epi(gw, K, accumulate=1)
            self.call("pose_gemm_bf16_tr", dy.data_ptr(), ldy, 1, x.data_ptr(), K, 1, N, K, M, self._splits(N, K, M),
                      C.byref(e))
        if bias is not None and bias.requires_grad:
            self.call("pose_colsum_bf16", dy.data_ptr(), M, N, ldy, self.flat.g32(bias, rows).data_ptr())
        if dx_name is None:
            return None
        dx = self.buf(dx_name, M, K)
        e = self._epi(dx, K, act=5 if act_u is not None else 0, residual=act_u, ldr=K, drop=act_drop)
        self.call("pose_gemm_bf16_tr", dy.data_ptr(), ldy, 0, w16.data_ptr(), K, 1, M, K, N, 1, C.byref(e))
        return dx

    def layernorm_bwd(self, x, dy, ln, M, dres, dx, rows=None, in_group=0, in_off=0):
        if rows is None:
            rows, in_group = M, M
        self.call("pose_layernorm_bwd_bf16", x.data_ptr(), dy.data_ptr(), self.flat.f32(ln.weight).data_ptr(),
                  float(ln.eps), M, rows, in_group, in_off, rows, 0, self.E,
                  dres.data_ptr() if dres is not None else None, dx.data_ptr(),
                  self.flat.g32(ln.weight).data_ptr() if ln.weight.requires_grad else None,
                  self.flat.g32(ln.bias).data_ptr() if ln.bias.requires_grad else None)
        return dx

    def attention_bwd(self, name, q, k, v, o, do, dq, dk, dv, Nq, Nk, heads, ldq, ldk, ldv, lddq, lddk, lddv, drop=None):
        E, B = self.E, self.B
        hd = E // heads
        lse = self.bufs[name + ".lse"]
        dws = self.buf(f"attn.D{heads}", B, heads, max(self.T_fin, self.T_img + 1), dtype=torch.float32)
        self.call("pose_attention_bwd_bf16", q, k, v, o.data_ptr(), do.data_ptr(), lse.data_ptr(), dq, dk, dv,
                  dws.data_ptr(), B, heads, Nq, Nk, hd, ldq, ldk, ldv, E, E, lddq, lddk, lddv, Nq * ldq, Nk * ldk,
                  Nk * ldv, Nq * E, Nq * E, Nq * lddq, Nk * lddk, Nk * lddv, 1.0 / math.sqrt(hd),
                  drop[0] if drop else 0.0, drop[1] if drop else 0)

    def encoder_block_bwd(self, pre, dx2, T, norm1, wqkv, bqkv, wo, bo, norm2, fc1, fc2, heads, p_drop=0.0, p_attn=0.0):
        """dx2 = gradient of the block output; returns the gradient of the block input (written in place)."""
        M, E, b = self.B * T, self.E, self.bufs
        dyf = self.drop_grad(f"d.m{M}", dx2, pre + "fc2", p_drop)
        du = self.linear_bwd(dyf, E, M, b[pre + "g"], fc2.weight, fc2.bias, dx_name=f"d.u{M}", act_u=b[pre + "g.u"],
                             act_drop=self.drop(pre + "act", p_drop))
        dh2 = self.linear_bwd(du, du.shape[1], M, b[pre + "h2"], fc1.weight, fc1.bias, dx_name=f"d.e{M}")
        dx1 = self.layernorm_bwd(b[pre + "x1"], dh2, norm2, M, dx2, dx2)
        dya = self.drop_grad(f"d.m{M}", dx1, pre + "proj", p_drop)
        do = self.linear_bwd(dya, E, M, b[pre + "o"], wo, bo, dx_name=f"d.e{M}")
        qkv = b[pre + "qkv"]
        dqkv = self.buf(f"d.qkv{M}", M, 3 * E)
        p, dp = qkv.data_ptr(), dqkv.data_ptr()
        self.attention_bwd(pre + "o", p, p + 2 * E, p + 4 * E, b[pre + "o"], do, dp, dp + 2 * E, dp + 4 * E, T, T, heads,
                           3 * E, 3 * E, 3 * E, 3 * E, 3 * E, 3 * E, drop=self.drop(pre + "attn", p_attn))
        dh = self.linear_bwd(dqkv, 3 * E, M, b[pre + "h"], wqkv, bqkv, dx_name=f"d.e{M}")
        x_in = b[self._block_input[pre]]
        return self.layernorm_bwd(x_in, dh, norm1, M, dx1, dx1)

    def mlp_residual_bwd(self, pre, dy, M, x_in, norm, mlp):
        b = self.bufs
        dyf = self.drop_grad(f"d.m{M}", dy, pre + "fc2", self.p_drop)
        du = self.linear_bwd(dyf, self.E, M, b[pre + "g"], mlp[3].weight, mlp[3].bias, dx_name=f"d.u{M}",
                             act_u=b[pre + "g.u"], act_drop=self.drop(pre + "act", self.p_drop))
        dh = self.linear_bwd(du, du.shape[1], M, b[pre + "h"], mlp[0].weight, mlp[0].bias, dx_name=f"d.e{M}")
        return self.layernorm_bwd(x_in, dh, norm, M, dy, dy)

    def cross_block_bwd(self, pre, blk, x_img, x_hm, d_img, d_hm):
        """(d_img, d_hm) = gradients of the block outputs, updated in place to the gradients of its inputs."""
        B, E, Ti, Th, b = self.B, self.E, self.T_img, self.T_hm, self.bufs
        heads = self.model.config.transformer_heads
        Mi, Mh = B * Ti, B * Th
        a1, a2 = blk.cross_attn_img_to_hm, blk.cross_attn_hm_to_img
        x_img1, x_hm1 = b[pre + "x_img1"], b[pre + "x_hm1"]
        self.mlp_residual_bwd(pre + "mh.", d_hm, Mh, x_hm1, blk.norm_hm_mlp, blk.mlp_hm)
        self.mlp_residual_bwd(pre + "mi.", d_img, Mi, x_img1, blk.norm_img_mlp, blk.mlp_img)
        # x_hm1 = x_hm + out_proj(attn(q2 = hm_q, kv2 = img_kv))
        dy2 = self.drop_grad(f"d.m{Mh}", d_hm, pre + "proj2", self.p_drop)
        do2 = self.linear_bwd(dy2, E, Mh, b[pre + "o2"], a2.out_proj.weight, a2.out_proj.bias, dx_name=f"d.e{Mh}")
        q2, kv2 = b[pre + "q2"], b[pre + "kv2"]
        dq2, dkv2 = self.buf(f"d.q{Mh}", Mh, E), self.buf(f"d.kv{Mi}", Mi, 2 * E)
        self.attention_bwd(pre + "o2", q2.data_ptr(), kv2.data_ptr(), kv2.data_ptr() + 2 * E, b[pre + "o2"], do2,
                           dq2.data_ptr(), dkv2.data_ptr(), dkv2.data_ptr() + 2 * E, Th, Ti, heads, E, 2 * E, 2 * E, E,
                           2 * E, 2 * E, drop=self.drop(pre + "attn2", self.p_attn))
        dhm_q = self.linear_bwd(dq2, E, Mh, b[pre + "hm_q"], a2.in_proj_weight, a2.in_proj_bias, rows=(0, E),
                                dx_name=f"d.e{Mh}")
        dimg_kv = self.linear_bwd(dkv2, 2 * E, Mi, b[pre + "img_kv"], a2.in_proj_weight, a2.in_proj_bias,
                                  rows=(E, 3 * E), dx_name=f"d.e{Mi}")
        self.layernorm_bwd(x_hm, dhm_q, blk.norm_hm_q, Mh, d_hm, d_hm)
        self.layernorm_bwd(x_img1, dimg_kv, blk.norm_img_kv, Mi, d_img, d_img)
        # x_img1 = x_img + out_proj(attn(q1 = img_q, kv1 = hm_kv))
        dy1 = self.drop_grad(f"d.m{Mi}", d_img, pre + "proj1", self.p_drop)
        do1 = self.linear_bwd(dy1, E, Mi, b[pre + "o1"], a1.out_proj.weight, a1.out_proj.bias, dx_name=f"d.e{Mi}")
        q1, kv1 = b[pre + "q1"], b[pre + "kv1"]
        dq1, dkv1 = self.buf(f"d.q{Mi}", Mi, E), self.buf(f"d.kv{Mh}", Mh, 2 * E)
        self.attention_bwd(pre + "o1", q1.data_ptr(), kv1.data_ptr(), kv1.data_ptr() + 2 * E, b[pre + "o1"], do1,
                           dq1.data_ptr(), dkv1.data_ptr(), dkv1.data_ptr() + 2 * E, Ti, Th, heads, E, 2 * E, 2 * E, E,
                           2 * E, 2 * E, drop=self.drop(pre + "attn1", self.p_attn))
        dimg_q = self.linear_bwd(dq1, E, Mi, b[pre + "img_q"], a1.in_proj_weight, a1.in_proj_bias, rows=(0, E),
                                 dx_name=f"d.e{Mi}")
        dhm_kv = self.linear_bwd(dkv1, 2 * E, Mh, b[pre + "hm_kv"], a1.in_proj_weight, a1.in_proj_bias,
                                 rows=(E, 3 * E), dx_name=f"d.e{Mh}")
        self.layernorm_bwd(x_img, dimg_q, blk.norm_img_q, Mi, d_img, d_img)
        self.layernorm_bwd(x_hm, dhm_kv, blk.norm_hm_kv, Mh, d_hm, d_hm)

    def backward(self, dout, section_done=None):
        """dout: gradient of the [B, J, 3] output (fp32).  Accumulates into the flat .grad buffer.
        section_done(flat, lo, hi) is called as soon as every kernel writing flat.grad[lo:hi) has been enqueued
        (the data-parallel trainer all-reduces that range while the rest of the backward runs)."""
        self._sp = _lib.stream_ptr()
        if not getattr(self, "saved", False):
            raise RuntimeError("backward needs a training-mode forward on this plan first")
        m, B, E, b = self.model, self.B, self.E, self.bufs
        c = m.config
        Ti, Th, Tf = self.T_img, self.T_hm, self.T_fin
        flat = self.flat
        if not flat.grads_attached():      # optimizer.zero_grad(set_to_none=True) dropped the views: new window
            flat.grad.zero_()
            flat.attach_grads()
        _lib.require_cuda(dout, "grad_output", torch.float32)
        hi = [flat.numel]

        def done(first_param):
            if section_done is not None:
                lo = flat.index[id(first_param)] if first_param is not None else 0
                section_done(flat, lo, hi[0])
                hi[0] = lo
        n_out = self.J * 3
        ld = (n_out + 7) // 8 * 8
        dy = self.buf("d.out", B, ld)
        self.call("pose_cast_f32_bf16_2d", dout.data_ptr(), n_out, B, n_out, dy.data_ptr(), ld)
        # ---- head ------------------------------------------------------------------------------------------
        lins = m.pose_head.linears()
        for i in range(len(lins) - 1, -1, -1):
            lin = lins[i]
            x = b[f"head{i - 1}"] if i > 0 else b["cls_out"]
            u = b[f"head{i - 1}.u"] if i > 0 else None
            dy = self.linear_bwd(dy, ld, B, x, lin.weight, lin.bias, dx_name=f"d.head{i}", act_u=u,
                                 act_drop=self.drop(f"head{i - 1}", self.p_head) if i > 0 else None)
            ld = dy.shape[1]
        dt = self.buf("d.fin", B * Tf, E, zero=True)
        self.layernorm_bwd(self.fin_out, dy, m.norm_out, B, None, dt, rows=1, in_group=Tf, in_off=0)
        done(m.norm_out.weight)
        # ---- final encoder ---------------------------------------------------------------------------------
        for i in range(len(m.final_encoder) - 1, -1, -1):
            blk = m.final_encoder[i]
            dt = self.encoder_block_bwd(f"fe{i}.", dt, Tf, blk.norm1, blk.attn.in_proj_weight, blk.attn.in_proj_bias,
                                        blk.attn.out_proj.weight, blk.attn.out_proj.bias, blk.norm2, blk.mlp[0],
                                        blk.mlp[3], c.transformer_heads, self.p_drop, self.p_attn)
            done(blk.norm1.weight)
        d = self.drop("final_pos_drop", self.p_drop)
        if d is not None:
            self.call("pose_dropout_bf16", dt.data_ptr(), dt.numel(), d[0], d[1], dt.data_ptr())
        self.call("pose_batch_rowsum_bf16", dt.data_ptr(), B, Tf, 0, Tf, E, flat.g32(m.final_pos_embed).data_ptr())
        self.call("pose_batch_rowsum_bf16", dt.data_ptr(), B, Tf, 0, 1, E, flat.g32(m.final_cls_token).data_ptr())
        d_img, d_hm = self.buf("d.img", B * Ti, E), self.buf("d.hm", B * Th, E)
        self.call("pose_token_slice_bf16", dt.data_ptr(), B, Tf, 1, Ti, E, d_img.data_ptr())
        self.call("pose_token_slice_bf16", dt.data_ptr(), B, Tf, 1 + Ti, Th, E, d_hm.data_ptr())
        # ---- cross-modal fusion ----------------------------------------------------------------------------
        n_cm = len(m.cross_modal_fusion_layers)
        for i in range(n_cm - 1, -1, -1):
            x_img = b[f"cm{i - 1}.mi.y"] if i > 0 else b["x_img"]
            x_hm = b[f"cm{i - 1}.mh.y"] if i > 0 else b["x_hm"]
            self.cross_block_bwd(f"cm{i}.", m.cross_modal_fusion_layers[i], x_img, x_hm, d_img, d_hm)
            done(m.cross_modal_fusion_layers[i].norm_img_q.weight)
        # ---- heat-map stream: pos_embed_hm and the patch embedding (key-points carry no gradient) -----------
        self.call("pose_batch_rowsum_bf16", d_hm.data_ptr(), B, Th, 0, Th, E, flat.g32(m.pos_embed_hm).data_ptr())
        hpe = m.heatmap_patch_embed.proj
        self.linear_bwd(d_hm, E, B * Th, b["phm"], hpe.weight, hpe.bias)
        done(hpe.weight)
        # ---- backbone --------------------------------------------------------------------------------------
        bb = m.vit_backbone
        dx = self.buf("d.bb", B * (Ti + 1), E, zero=True)
        self.layernorm_bwd(self.bb_out, d_img, bb.norm, B * Ti, None, dx, rows=Ti, in_group=Ti + 1, in_off=1)
        for i in range(len(bb.blocks) - 1, -1, -1):
            blk = bb.blocks[i]
            dx = self.encoder_block_bwd(f"bb{i}.", dx, Ti + 1, blk.norm1, blk.attn.qkv.weight, blk.attn.qkv.bias,
                                        blk.attn.proj.weight, blk.attn.proj.bias, blk.norm2, blk.mlp.fc1, blk.mlp.fc2,
                                        bb.num_heads)
            if i % 3 == 0 or i < 3:      # groups of three; the last blocks one by one: a short exposed all-reduce tail
                done(blk.norm1.weight)
        if bb.pos_embed.requires_grad:
            self.call("pose_batch_rowsum_bf16", dx.data_ptr(), B, Ti + 1, 0, Ti + 1, E, flat.g32(bb.pos_embed).data_ptr())
        if bb.cls_token.requires_grad:
            self.call("pose_batch_rowsum_bf16", dx.data_ptr(), B, Ti + 1, 0, 1, E, flat.g32(bb.cls_token).data_ptr())
        dtok = self.buf("d.tok", B * Ti, E)
        self.call("pose_token_slice_bf16", dx.data_ptr(), B, Ti + 1, 1, Ti, E, dtok.data_ptr())
        self.linear_bwd(dtok, E, B * Ti, b["pimg"], bb.patch_embed.proj.weight, bb.patch_embed.proj.bias)
        done(None)

    @property
    def _block_input(self):
        """name of the buffer holding each encoder block's input (the previous block's output)."""
        tbl = getattr(self, "_block_input_tbl", None)
        if tbl is None:
            tbl = {}
            n_bb = len(self.model.vit_backbone.blocks)
            for i in range(n_bb):
                tbl[f"bb{i}."] = "bb.x0" if i == 0 else f"bb{i - 1}.x2"
            for i in range(len(self.model.final_encoder)):
                tbl[f"fe{i}."] = "fin.x0" if i == 0 else f"fe{i - 1}.x2"
            self._block_input_tbl = tbl
        return tbl
