"""Training-mode launch plan of CNNPoseEstimation (reference: src/models/cnn.py:641-665 in ``model.train()``, and
``loss.backward()`` over it: src/train.py:83-92).

Forward: every ConvBnAct is conv (tcgen05 GEMM / implicit GEMM, bf16) -> per-channel batch statistics -> normalise +
activation (+ residual add, + write into a channel slice of a concatenation); BatchNorm uses batch statistics and
updates the running buffers exactly like nn.BatchNorm2d.  Backward mirrors the model block by block: BatchNorm
backward (two passes), weight gradients and data gradients as tcgen05 GEMMs that read the saved activations and the
weights in place (1x1 convolutions), as an implicit GEMM over 64-pixel patches (spatial convolutions, weight
gradient) or as the forward convolution kernel with flipped weights (spatial convolutions, data gradient).
Parameter gradients accumulate in the flat fp32 ``.grad`` buffer (params.FlatParams).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..params import FlatParams
from ..utils import split_k, wgrad_splits, activation_id


def _ceil8(n):
    return (n + 7) // 8 * 8


# GEMM epilogue id "multiply by the activation derivative the forward GEMM saved" (csrc/gemm_tcgen05.cu); the table is kept
# so that call sites read as act -> act'
ACT_GRAD = {1: 5, 2: 5, 3: 5}


class CnnTrainPlan:
    def __init__(self, model, B, device):
        from . import cnn as cm
        self.cm = cm
        self.model, self.B, self.dev = model, B, device
        self.lib = _lib.lib()
        self._sp = None          # stream handle of the current forward / backward (see call)
        c = model.config
        S = int(c.heatmap_size)
        if tuple(c.image_size) != (S, S):
            raise ValueError("CNNPoseEstimation needs heatmap_size == image height == width (cnn.py:648)")
        if c.in_channels != 4 + c.num_joints or c.in_channels > 32:
            raise NotImplementedError("conv1 operand packs 3 RGB + 1 depth + J heat-maps into 32 channels")
        if c.residual_scale != 1.0:
            raise NotImplementedError("training path: residual_scale must be 1.0 (the reference default)")
        self.S, self.J = S, c.num_joints
        self.act = activation_id(c.activation)
        if self.act not in (1, 2):
            raise NotImplementedError("CNN training activation: silu or relu")
        self.flat = FlatParams.of(model.parameters())
        self.bufs, self.rec = {}, {}
        self.launches = 0
        self.step_count = 0
        self.drop_p = float(c.regression_dropout)
        # BatchNorm backward: reduction pass carried by the producer of the incoming gradient (A/B switches).  Measured on
        # the 768-channel layers at B = 128: SE-gate backward + tail 123 us (74 without) but the BatchNorm backward drops from
        # 182 to 109 us: on.  The depthwise data gradient is issue bound already (100 us); with the tail it runs two CTAs per
        # SM at twice the instructions (294 us) against 100 + 75 us for the separate reduction pass: off.
        self.fuse_bn_bwd = os.environ.get("POSE_FUSE_BN_BWD", "1") != "0"
        self.fuse_bn_bwd_dw = os.environ.get("POSE_FUSE_BN_BWD_DW", "0") == "1"
        self._emit_parts = None
        # fp32 workspace zeroed once per step: BN sums, gate gradients, spatial-conv weight-gradient staging
        self._ws_off, self._ws_items = 0, {}
        self._pk32_off, self._pk16_off = 0, 0
        self._repack, self._unpack = [], []
        self.wsbuf = None
        self._layout()

    # ---- static layout: packed parameter copies, workspace slots ---------------------------------------
    def _ws(self, name, n):
        off = self._ws_items.get(name)
        if off is None:
            off = self._ws_off
            self._ws_items[name] = off
            self._ws_off += (n + 63) // 64 * 64
        return off

    def _layout(self):
        cm, m, flat = self.cm, self.model, self.flat
        self.pk = {}
        for name, mod in m.named_modules():
            if isinstance(mod, cm.ConvBnAct):
                conv = mod.conv
                co, ci, kh, kw = conv.weight.shape
                src = flat.index[id(conv.weight)]
                if conv.groups > 1:
                    n9 = (9 * co + 63) // 64 * 64
                    self.pk[name] = ("dw", self._pk32_off, self._pk32_off + n9)
                    self._repack.append((src, self._pk32_off, 0, co, 0, 0, 0))
                    self._repack.append((src, self._pk32_off + n9, 6, co, 0, 0, 0))     # flipped taps: data gradient
                    self._pk32_off += 2 * n9
                elif kh > 1 or conv.stride[0] > 1:
                    cp = 32 if ci <= 32 else (ci + 63) // 64 * 64     # the 21-channel network input: 64-byte pixels
                    fwd = self._pk16_off
                    self._pk16_off += (co * kh * kw * cp + 63) // 64 * 64
                    bwd = None
                    if name != "conv1.0" and conv.stride[0] == 1:     # data gradient by the flipped forward conv
                        bwd = self._pk16_off
                        self._pk16_off += (ci * kh * kw * co + 63) // 64 * 64
                        self._repack.append((src, bwd, 2, co, ci, kh, cp))
                    self._repack.append((src, fwd, 1, co, ci, kh, cp))
                    stage = self._ws(name + ".dwk", co * kh * kw * cp)
                    self._unpack.append((stage, src, 5, co, ci, kh, cp))
                    self.pk[name] = ("conv", fwd, bwd, cp, stage, len(self._unpack) - 1)
            elif isinstance(mod, cm.SEBlock):
                w2 = mod.fc[2].weight
                cr = w2.shape[1]
                if cr % 8:
                    self.pk[name] = ("se", self._pk16_off)
                    self._repack.append((flat.index[id(w2)], self._pk16_off, 3, w2.shape[0], cr, _ceil8(cr), 0))
                    self._pk16_off += (w2.shape[0] * _ceil8(cr) + 63) // 64 * 64
        self.pk32 = torch.zeros(max(self._pk32_off, 64), dtype=torch.float32, device=self.dev)
        self.pk16 = torch.zeros(max(self._pk16_off, 64), dtype=torch.bfloat16, device=self.dev)

        def table(entries):
            arr = (_lib.PoseRepackEntry * len(entries))()
            for i, (src, dst, kind, d0, d1, d2, d3) in enumerate(entries):
                arr[i].src, arr[i].dst, arr[i].kind = src, dst, kind
                arr[i].d0, arr[i].d1, arr[i].d2, arr[i].d3 = d0, d1, d2, d3
            raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
            return torch.from_numpy(raw).to(self.dev)
        self.repack_table, self.n_repack = table(self._repack), len(self._repack)
        self.unpack_table, self.n_unpack = table(self._unpack), len(self._unpack)
        self._pk_version = None
        self.bn_tracked = [mod.num_batches_tracked for mod in m.modules() if isinstance(mod, nn.BatchNorm2d)]

    # ---- helpers ---------------------------------------------------------------------------------------
    def buf(self, name, *shape, dtype=torch.bfloat16):
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=self.dev)
            self.bufs[name] = t
        return t

    def _seed(self, site):
        """Dropout seed of a head layer: the step count is part of it unless a pose_step_state is bound (then the per-step
        part lives in device memory and the seed is constant, so the captured step can be replayed)."""
        return site + 1 if _lib.step_state_bound() else self.step_count * 16 + site

    def call(self, name, *args):
        # the stream handle is looked up once per forward / backward (a property chain in torch, ~3 us: it was the largest
        # single item of the per-launch host cost of a step with 500-600 launches)
        rc = getattr(self.lib, name)(*args, self._sp if self._sp is not None else _lib.stream_ptr())
        if rc:
            _lib.check(rc, name)
        self.launches += 1

    def ws(self, name, n):
        off = self._ws(name, n)
        if self.wsbuf is None or off + n > self.wsbuf.numel():
            raise RuntimeError("workspace layout changed after allocation")
        return self.wsbuf[off:off + n]

    def _epi(self, out, ldc, bias=None, act=0, residual=None, ldr=0, preact=None, accumulate=0, col_off=0):
        e = _lib.PoseGemmEpilogue()
        e.bias = bias.data_ptr() if bias is not None else None
        e.residual = residual if isinstance(residual, int) or residual is None else residual.data_ptr()
        esz = 4 if out.dtype == torch.float32 else 2
        e.C = out.data_ptr() + col_off * esz
        e.ldc, e.ldr, e.act = ldc, ldr, act
        e.out_dtype = 0 if out.dtype == torch.float32 else 1
        e.out_scale, e.res_scale = 1.0, 1.0
        e.preact = preact.data_ptr() if preact is not None else None
        e.accumulate = accumulate
        return e

    def gemm(self, a_ptr, lda, w_ptr, ldw, M, N, K, e):
        self.call("pose_gemm_bf16_ex", a_ptr, lda, w_ptr, ldw, M, N, K, C.byref(e))

    def gemm_tr(self, a_ptr, lda, a_mn, w_ptr, ldw, b_mn, M, N, K, e, splits=1):
        self.call("pose_gemm_bf16_tr", a_ptr, lda, a_mn, w_ptr, ldw, b_mn, M, N, K, splits, C.byref(e))

    @staticmethod
    def _splits(n_out, n_in, rows):
        return wgrad_splits(n_out, n_in, rows)

    # ---- ConvBnAct ------------------------------------------------------------------------------------
    def cba_fwd(self, name, cba, x, shape, act="default", out=None, ld_out=None, col_off=0, residual=None, ld_res=0,
                pool=None):
        """x: [B,H,W,Cin] bf16 (channels-last).  Returns (a, (B,Ho,Wo,Cout)); records what the backward needs."""
        Bn, H, W, cin = shape
        conv, bn = cba.conv, cba.norm
        co = conv.out_channels
        kh = conv.kernel_size[0]
        stride, dil, pad = conv.stride[0], conv.dilation[0], conv.padding[0]
        a_id = 0 if (act is None or cba.activation is None) else (self.act if act == "default" else activation_id(act))
        flat = self.flat
        if conv.groups > 1:
            Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
            y = self.buf(name + ".y", Bn * Ho * Wo, co)
            wd = self.pk32[self.pk[name][1]:]
            # the depthwise kernel emits the first stage of the batch statistics of its own output (no separate pass)
            part = self.partials()
            self.call("pose_dwconv3x3_bn_stats_bf16", x.data_ptr(), Bn, H, W, co, wd.data_ptr(), stride, y.data_ptr(),
                      part.data_ptr(), part.numel())
            stat_parts = Bn * self.lib.pose_dwconv3x3_pool_parts(H, W, stride)
            kind = "dw"
        else:
            spatial = name in self.pk
            if spatial:
                _, fwd, bwd, cp, stage, _ui = self.pk[name]
                Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
                Wo = (W + 2 * pad - dil * (kh - 1) - 1) // stride + 1
            else:
                Ho, Wo = H, W
            M = Bn * Ho * Wo
            y = self.buf(name + ".y", M, co)
            mr = self.buf(name + ".mr", 2 * co, dtype=torch.float32)
            ss = self.buf(name + ".ss", 2 * co, dtype=torch.float32)
            e = self._epi(y, co)
            fused = co % 32 == 0
            if fused:
                # batch statistics from the fp32 accumulators in the epilogue + the fold, in the same call (pose_bn_fuse)
                part = self.partials()
                f = _lib.PoseBnFuse()
                f.partials, f.cap_floats = part.data_ptr(), part.numel()
                f.gamma, f.beta = flat.f32(bn.weight).data_ptr(), flat.f32(bn.bias).data_ptr()
                f.eps, f.momentum, f.count = float(bn.eps), float(bn.momentum), M
                f.mean_rstd, f.scale_shift = mr.data_ptr(), ss.data_ptr()
                f.running_mean, f.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
                e.bn = C.pointer(f)
            # a dilated 'same' convolution whose dilation reaches across the whole map (WASP rate 18 on the 16 x 16 map of a
            # 256 x 256 input): every tap but the centre reads only zero padding, so the layer IS the 1x1 convolution with
            # the centre-tap weights -- a plain GEMM over the packed KRSC weights with pitch K*K*Cp (1/9 of the work)
            center = spatial and kh > 1 and stride == 1 and 2 * pad == dil * (kh - 1) and dil >= H and dil >= W and cin == cp
            if center:
                ctr = (kh // 2 * kh + kh // 2) * cp
                self.gemm(x.data_ptr(), cin, self.pk16[fwd + ctr:].data_ptr(), kh * kh * cp, M, co, cin, e)
                kind = "conv"
            elif spatial:
                self.call("pose_conv2d_bf16", x.data_ptr(), Bn, H, W, cin, self.pk16[fwd:].data_ptr(), co, kh, kh, stride,
                          dil, pad, C.byref(e))
                kind = "conv"
            else:
                w16 = flat.w16(conv.weight)
                self.gemm(x.data_ptr(), cin, w16.data_ptr(), cin, M, co, cin, e)
                kind = "1x1"
            if fused:
                self.launches += 1          # the fold launched by the same call
        M = Bn * Ho * Wo
        part = self.partials()
        if kind == "dw":
            mr = self.buf(name + ".mr", 2 * co, dtype=torch.float32)
            ss = self.buf(name + ".ss", 2 * co, dtype=torch.float32)
            self.call("pose_bn_finalize_parts", part.data_ptr(), stat_parts, M, flat.f32(bn.weight).data_ptr(),
                      flat.f32(bn.bias).data_ptr(), float(bn.eps), float(bn.momentum), co, mr.data_ptr(), ss.data_ptr(),
                      bn.running_mean.data_ptr(), bn.running_var.data_ptr())
        elif not fused:
            self.call("pose_bn_stats_bf16", y.data_ptr(), M, co, co, part.data_ptr(), part.numel())
            self.call("pose_bn_finalize", part.data_ptr(), part.numel(), M, flat.f32(bn.weight).data_ptr(),
                      flat.f32(bn.bias).data_ptr(), float(bn.eps), float(bn.momentum), co, mr.data_ptr(), ss.data_ptr(),
                      bn.running_mean.data_ptr(), bn.running_var.data_ptr())
        if out is None:
            out = self.buf(name + ".a", M, co)
            ld_out = co
        optr = out.data_ptr() + 2 * col_off
        self._pool = None
        if pool is not None and residual is None and ld_out == co and col_off == 0:
            # the layer feeds an SE / ECA block: its squeeze (per-image channel sums) rides along with the normalisation
            HWo = Ho * Wo
            parts = max(1, min(HWo // 8 if HWo >= 8 else 1, (148 * 4) // Bn))      # one wave of resident CTAs
            pl = self.buf(pool, Bn, parts, co, dtype=torch.float32)
            self.call("pose_bn_apply_pool_bf16", y.data_ptr(), Bn, HWo, co, ss.data_ptr(), a_id, optr, pl.data_ptr(), parts)
            self._pool = pl
        else:
            self.call("pose_bn_apply_bf16", y.data_ptr(), M, co, ss.data_ptr(), a_id, 1.0,
                      residual.data_ptr() if residual is not None else None, ld_res, optr, ld_out)
        self.rec[name] = dict(kind=kind, x=x, shape=shape, y=y, mr=mr, ss=ss, act=a_id, M=M, co=co, cin=cin,
                              oshape=(Bn, Ho, Wo, co), cba=cba, center=kind == "conv" and center)
        return out, (Bn, Ho, Wo, co)

    def partials(self):
        """Scratch for the two-stage per-channel reductions (stream-ordered reuse by every BatchNorm)."""
        t = self.bufs.get("bn.partials")
        if t is None:
            t = torch.empty(8 << 20, dtype=torch.float32, device=self.dev)
            self.bufs["bn.partials"] = t
            self.bufs["bn.coef"] = torch.empty(2 * 3072, dtype=torch.float32, device=self.dev)
        return t

    def zero_bias(self, n):
        t = self.bufs.get("zero_bias")
        if t is None or t.numel() < n:
            t = torch.zeros(max(n, 4096), dtype=torch.float32, device=self.dev)
            self.bufs["zero_bias"] = t
        return t

    def cba_bwd(self, name, dA, ld_da, need_dx=True, dx_add=None, ld_add=0, da_off=0, dz_parts=None, bn_next=None):
        """dA: gradient of the layer output (pitch ld_da, column offset da_off).  Accumulates dgamma / dbeta / dW.
        Returns dx [M_in, Cin] (+ dx_add fused into the data-gradient epilogue) or None.

        dz_parts: the producer of `dA` already multiplied by act'(z) of THIS layer and left `dz_parts` rows of partial sums in
        the shared scratch (the reduction pass of the BatchNorm backward is skipped).  bn_next: record of the layer whose
        output this layer consumed; where the data gradient is produced by a kernel that can carry that layer's reduction
        (stride-1 depthwise), it does, and self._emit_parts tells the caller."""
        r = self.rec[name]
        cba, flat = r["cba"], self.flat
        conv, bn = cba.conv, cba.norm
        co, cin, M = r["co"], r["cin"], r["M"]
        Bn, H, W, _ = r["shape"]
        _, Ho, Wo, _ = r["oshape"]
        dy = self.buf(name + ".dy", M, co)
        part = self.partials()
        self._emit_parts = None
        if dz_parts is not None:
            self.call("pose_bn_bwd_from_dz_bf16", dA.data_ptr() + 2 * da_off, ld_da, r["y"].data_ptr(), M, co, r["ss"].data_ptr(),
                      r["mr"].data_ptr(), part.data_ptr(), dz_parts, self.bufs["bn.coef"].data_ptr(), dy.data_ptr(),
                      flat.g32(bn.weight).data_ptr(), flat.g32(bn.bias).data_ptr())
        else:
            self.call("pose_bn_bwd_bf16", dA.data_ptr() + 2 * da_off, ld_da, r["y"].data_ptr(), M, co, r["ss"].data_ptr(),
                      r["mr"].data_ptr(), r["act"], 1.0, part.data_ptr(), part.numel(), self.bufs["bn.coef"].data_ptr(),
                      dy.data_ptr(), flat.g32(bn.weight).data_ptr(), flat.g32(bn.bias).data_ptr())
        x = r["x"]
        kh = conv.kernel_size[0]
        stride, dil, pad = conv.stride[0], conv.dilation[0], conv.padding[0]
        add_ptr = dx_add.data_ptr() if dx_add is not None else None
        if r["kind"] == "dw":
            wd = self.pk32[self.pk[name][1]:]
            if dx_add is not None and ld_add not in (0, co):
                raise RuntimeError("depthwise dx_add must be compact")
            dx = self.buf(name + ".dx", Bn * H * W, co) if need_dx else None
            if need_dx and stride == 1:
                # stride 1: the data gradient is the forward (shared-memory tiled) kernel over dY with flipped taps
                wf = self.pk32[self.pk[name][2]:]
                fuse = (bn_next is not None and dx_add is None and self.fuse_bn_bwd_dw and bn_next["act"] in (1, 2)
                        and bn_next["co"] == co and bn_next["M"] == Bn * H * W)
                if fuse:
                    # ... which also carries the reduction pass of the BatchNorm in front of this layer (emits dz)
                    parts = Bn * self.lib.pose_dwconv3x3_pool_parts(Ho, Wo, 1)
                    fuse = parts * 2 * co <= part.numel()
                if fuse:
                    self.call("pose_dwconv3x3_bnbwd_bf16", dy.data_ptr(), Bn, Ho, Wo, co, wf.data_ptr(), bn_next["y"].data_ptr(),
                              bn_next["ss"].data_ptr(), bn_next["act"], dx.data_ptr(), part.data_ptr(), part.numel())
                    self.call("pose_dwconv3x3_bwd_bf16", dy.data_ptr(), x.data_ptr(), wd.data_ptr(), Bn, H, W, co, stride, None,
                              None, flat.g32(conv.weight).data_ptr())
                    self._emit_parts = parts
                    return dx
                self.call("pose_dwconv3x3_bf16", dy.data_ptr(), Bn, Ho, Wo, co, wf.data_ptr(), self.zero_bias(co).data_ptr(), 1,
                          0, dx.data_ptr(), None, 0)
                if dx_add is not None:
                    self.call("pose_add_bf16", dx.data_ptr(), add_ptr, dx.numel(), dx.data_ptr())
                self.call("pose_dwconv3x3_bwd_bf16", dy.data_ptr(), x.data_ptr(), wd.data_ptr(), Bn, H, W, co, stride, None,
                          None, flat.g32(conv.weight).data_ptr())
            else:
                self.call("pose_dwconv3x3_bwd_bf16", dy.data_ptr(), x.data_ptr(), wd.data_ptr(), Bn, H, W, co, stride, add_ptr,
                          dx.data_ptr() if need_dx else None, flat.g32(conv.weight).data_ptr())
            return dx
        if r["kind"] == "conv":
            _, fwd, bwd, cp, stage, ui = self.pk[name]
            st = self.wsbuf[stage:]
            if r["center"]:
                # centre-tap-only layer (see cba_fwd): both gradients are the 1x1 convolution's GEMMs on strided views of
                # the KRSC staging buffer / the flipped [Ci,K,K,Co] weights
                ctr = kh // 2 * kh + kh // 2
                self.gemm_tr(dy.data_ptr(), co, 1, x.data_ptr(), cin, 1, co, cin, M,
                             self._epi(st[ctr * cp:], kh * kh * cp, accumulate=1), self._splits(co, cin, M))
                self.call("pose_param_repack", self.unpack_table.data_ptr() + ui * C.sizeof(_lib.PoseRepackEntry), 1,
                          self.wsbuf.data_ptr(), flat.grad.data_ptr(), None)
                if not need_dx:
                    return None
                dx = self.buf(name + ".dx", Bn * H * W, cin)
                e = self._epi(dx, cin, residual=add_ptr, ldr=ld_add or cin)
                self.gemm(dy.data_ptr(), co, self.pk16[bwd + ctr * co:].data_ptr(), kh * kh * co, M, cin, co, e)
                return dx
            tiles = ((co + 127) // 128) * ((kh * kh * cp + 127) // 128)
            splits = split_k(tiles, (M + 63) // 64)
            self.call("pose_conv2d_wgrad_bf16", dy.data_ptr(), x.data_ptr(), Bn, H, W, r["cin"], co, kh, kh, stride, dil, pad,
                      st.data_ptr(), splits)
            # the gradient was staged in KRSC order: add it into the [Co,Ci,K,K] .grad view
            self.call("pose_param_repack", self.unpack_table.data_ptr() + ui * C.sizeof(_lib.PoseRepackEntry), 1,
                      self.wsbuf.data_ptr(), flat.grad.data_ptr(), None)
            if not need_dx:
                return None
            if bwd is not None:
                # stride-1 'same' convolution: dx = conv(dy, flipped / transposed weights), same dilation and padding
                dx = self.buf(name + ".dx", Bn * H * W, cin)
                e = self._epi(dx, cin, residual=add_ptr, ldr=ld_add or cin)
                self.call("pose_conv2d_bf16", dy.data_ptr(), Bn, Ho, Wo, co, self.pk16[bwd:].data_ptr(), cin, kh, kh, 1, dil,
                          pad, C.byref(e))
                return dx
            if kh != 1:
                raise NotImplementedError("data gradient of a strided spatial convolution")
            # strided 1x1: compact data gradient, scattered (added) into the full-resolution gradient
            dxs = self.buf(name + ".dxs", M, cin)
            w16 = self.pk16[fwd:]
            self.gemm_tr(dy.data_ptr(), co, 0, w16.data_ptr(), cp, 1, M, cin, co, self._epi(dxs, cin))
            if dx_add is None or ld_add not in (0, cin):
                raise RuntimeError("strided 1x1 data gradient is added into an existing compact gradient")
            self.call("pose_scatter_strided_add_bf16", dxs.data_ptr(), Bn, Ho, Wo, H, W, cin, stride, dx_add.data_ptr())
            return dx_add
        # 1x1 stride 1
        w16 = flat.w16(conv.weight)
        gw = flat.g32(conv.weight)
        self.gemm_tr(dy.data_ptr(), co, 1, x.data_ptr(), cin, 1, co, cin, M, self._epi(gw, cin, accumulate=1),
                     self._splits(co, cin, M))
        if not need_dx:
            return None
        dx = self.buf(name + ".dx", M, cin)
        self.gemm_tr(dy.data_ptr(), co, 0, w16.data_ptr(), cin, 1, M, cin, co,
                     self._epi(dx, cin, residual=add_ptr, ldr=ld_add or cin))
        return dx

    # ---- attention blocks -----------------------------------------------------------------------------
    def pool_sums(self, name, x, Bn, HW, ch):
        parts = max(1, min(HW // 8 if HW >= 8 else 1, (148 * 4 + Bn - 1) // Bn))
        pool = self.buf(name, Bn, parts, ch, dtype=torch.float32)
        self.call("pose_pool_sum_bf16", x.data_ptr(), Bn, HW, ch, pool.data_ptr(), parts)
        return pool

    def att_fwd(self, name, att, x, shape, pool=None):
        cm, flat = self.cm, self.flat
        Bn, H, W, ch = shape
        HW = H * W
        if isinstance(att, (cm.SEBlock, cm.ECABlock)):
            if pool is None:
                pool = self.pool_sums(name + ".pool", x, Bn, HW, ch)
            gate = self.buf(name + ".gate", Bn, ch, dtype=torch.float32)
            if isinstance(att, cm.SEBlock):
                w1, w2 = att.fc[0].weight, att.fc[2].weight
                cr = w1.shape[0]
                ldp = _ceil8(cr)
                mean16 = self.buf(name + ".mean", Bn, ch)
                self.call("pose_sums_to_bf16", pool.data_ptr(), pool.shape[1], Bn, ch, 1.0 / HW, mean16.data_ptr())
                hid, u1 = self.buf(name + ".hid", Bn, ldp), self.buf(name + ".u1", Bn, ldp)
                e = self._epi(hid, ldp, act=activation_id(att.activation), preact=u1)
                self.gemm(mean16.data_ptr(), ch, flat.w16(w1).data_ptr(), ch, Bn, cr, ch, e)
                w2p = self.pk16[self.pk[name][1]:] if name in self.pk else flat.w16(w2)
                self.gemm(hid.data_ptr(), ldp, w2p.data_ptr(), ldp, Bn, ch, cr, self._epi(gate, ch, act=4))
            else:
                k = att.conv.weight.shape[-1]
                self.call("pose_eca_gate", pool.data_ptr(), pool.shape[1], 1.0 / HW, flat.f32(att.conv.weight).data_ptr(), k,
                          Bn, ch, gate.data_ptr(), None)
            out = self.buf(name + ".out", Bn * HW, ch)
            self.call("pose_channel_affine_bf16", x.data_ptr(), gate.data_ptr(), None, Bn, HW, ch, out.data_ptr())
            self.rec[name] = dict(x=x, shape=shape, pool=pool, gate=gate, att=att)
            return out
        # CoordAttention (cnn.py:48-98): the shared 1x1 conv + BatchNorm runs over the B*(H+W) pooled positions
        mid = att.conv1.out_channels
        rows = Bn * (H + W)
        P = self.buf(name + ".P", rows, ch)
        self.call("pose_coord_pool_bf16", x.data_ptr(), Bn, H, W, ch, P.data_ptr())
        y1 = self.buf(name + ".y1", rows, mid)
        e = self._epi(y1, mid, bias=flat.f32(att.conv1.bias))
        self.gemm(P.data_ptr(), ch, flat.w16(att.conv1.weight).data_ptr(), ch, rows, mid, ch, e)
        bn = att.bn1
        part = self.partials()
        self.call("pose_bn_stats_bf16", y1.data_ptr(), rows, mid, mid, part.data_ptr(), part.numel())
        mr = self.buf(name + ".mr", 2 * mid, dtype=torch.float32)
        ss = self.buf(name + ".ss", 2 * mid, dtype=torch.float32)
        self.call("pose_bn_finalize", part.data_ptr(), part.numel(), rows, flat.f32(bn.weight).data_ptr(), flat.f32(bn.bias).data_ptr(),
                  float(bn.eps), float(bn.momentum), mid, mr.data_ptr(), ss.data_ptr(), bn.running_mean.data_ptr(),
                  bn.running_var.data_ptr())
        a1 = self.buf(name + ".a1", rows, mid)
        self.call("pose_bn_apply_bf16", y1.data_ptr(), rows, mid, ss.data_ptr(), 2, 1.0, None, 0, a1.data_ptr(), mid)
        G = self.buf(name + ".G", rows, 2 * ch)
        for j, cv in enumerate((att.conv_h, att.conv_w)):
            e = self._epi(G, 2 * ch, bias=flat.f32(cv.bias), act=4, col_off=j * ch)
            self.gemm(a1.data_ptr(), mid, flat.w16(cv.weight).data_ptr(), mid, rows, ch, mid, e)
        out = self.buf(name + ".out", Bn * HW, ch)
        self.call("pose_coord_apply_bf16", x.data_ptr(), G.data_ptr(), Bn, H, W, ch, out.data_ptr())
        self.rec[name] = dict(x=x, shape=shape, P=P, y1=y1, mr=mr, ss=ss, a1=a1, G=G, att=att)
        return out

    def att_bwd(self, name, dout, add=None, bn_next=None):
        """dout: gradient of the gated output [B*HW, C]; returns the gradient of the block input (+ add).  bn_next: as in
        cba_bwd (the SE / ECA gate backward can carry the reduction pass of the BatchNorm in front of the block)."""
        cm, flat = self.cm, self.flat
        r = self.rec[name]
        self._emit_parts = None
        att, x = r["att"], r["x"]
        Bn, H, W, ch = r["shape"]
        HW = H * W
        dx = self.buf(name + ".dx", Bn * HW, ch)
        addp = add.data_ptr() if add is not None else None
        if isinstance(att, (cm.SEBlock, cm.ECABlock)):
            gate, pool = r["gate"], r["pool"]
            dgate = self.ws(name + ".dgate", Bn * ch)
            self.call("pose_gate_bwd_reduce_bf16", dout.data_ptr(), x.data_ptr(), Bn, HW, ch, dgate.data_ptr())
            dmean = self.buf(name + ".dmean", Bn, ch)
            if isinstance(att, cm.SEBlock):
                w1, w2 = att.fc[0].weight, att.fc[2].weight
                cr = w1.shape[0]
                ldp = _ceil8(cr)
                b = self.bufs
                dz2 = self.buf(name + ".dz2", Bn, ch)
                self.call("pose_sigmoid_bwd", dgate.data_ptr(), gate.data_ptr(), Bn * ch, dz2.data_ptr())
                hid, u1, mean16 = b[name + ".hid"], b[name + ".u1"], b[name + ".mean"]
                self.gemm_tr(dz2.data_ptr(), ch, 1, hid.data_ptr(), ldp, 1, ch, cr, Bn, self._epi(flat.g32(w2), cr, accumulate=1))
                w2p = self.pk16[self.pk[name][1]:] if name in self.pk else flat.w16(w2)
                du1 = self.buf(name + ".du1", Bn, ldp)
                e = self._epi(du1, ldp, act=ACT_GRAD[activation_id(att.activation)], residual=u1, ldr=ldp)
                self.gemm_tr(dz2.data_ptr(), ch, 0, w2p.data_ptr(), ldp, 1, Bn, cr, ch, e)
                self.gemm_tr(du1.data_ptr(), ldp, 1, mean16.data_ptr(), ch, 1, cr, ch, Bn,
                             self._epi(flat.g32(w1), ch, accumulate=1))
                self.gemm_tr(du1.data_ptr(), ldp, 0, flat.w16(w1).data_ptr(), ch, 1, Bn, ch, cr, self._epi(dmean, ch))
            else:
                k = att.conv.weight.shape[-1]
                self.call("pose_eca_bwd", dgate.data_ptr(), None, gate.data_ptr(), pool.data_ptr(), pool.shape[1], 1.0 / HW,
                          flat.f32(att.conv.weight).data_ptr(), k, Bn, ch, 0, dmean.data_ptr(),
                          flat.g32(att.conv.weight).data_ptr())
            if (bn_next is not None and add is None and self.fuse_bn_bwd and bn_next["act"] in (1, 2) and bn_next["co"] == ch
                    and bn_next["M"] == Bn * HW):
                part = self.partials()
                n_parts = C.c_int(0)
                self.call("pose_gate_bwd_apply_bn_bf16", dout.data_ptr(), gate.data_ptr(), dmean.data_ptr(), 1.0 / HW, Bn, HW, ch,
                          bn_next["y"].data_ptr(), bn_next["ss"].data_ptr(), bn_next["act"], dx.data_ptr(), part.data_ptr(),
                          part.numel(), C.byref(n_parts))
                self._emit_parts = n_parts.value
                return dx
            self.call("pose_gate_bwd_apply_bf16", dout.data_ptr(), gate.data_ptr(), dmean.data_ptr(), 1.0 / HW, Bn, HW, ch, addp,
                      dx.data_ptr())
            return dx
        mid = att.conv1.out_channels
        rows = Bn * (H + W)
        G, a1, y1, P = r["G"], r["a1"], r["y1"], r["P"]
        dZ = self.buf(name + ".dZ", rows, 2 * ch)
        self.call("pose_coord_bwd_reduce_bf16", dout.data_ptr(), x.data_ptr(), G.data_ptr(), Bn, H, W, ch, dZ.data_ptr())
        da1 = self.buf(name + ".da1", rows, mid)
        for j, cv in enumerate((att.conv_h, att.conv_w)):
            zp = dZ.data_ptr() + 2 * j * ch
            self.gemm_tr(zp, 2 * ch, 1, a1.data_ptr(), mid, 1, ch, mid, rows,
                         self._epi(flat.g32(cv.weight).view(ch, mid), mid, accumulate=1), self._splits(ch, mid, rows))
            self.call("pose_colsum_bf16", zp, rows, ch, 2 * ch, flat.g32(cv.bias).data_ptr())
            e = self._epi(da1, mid, residual=da1 if j == 1 else None, ldr=mid)
            self.gemm_tr(zp, 2 * ch, 0, flat.w16(cv.weight).data_ptr(), mid, 1, rows, mid, ch, e)
        bn = att.bn1
        dy1 = self.buf(name + ".dy1", rows, mid)
        part = self.partials()
        self.call("pose_bn_bwd_bf16", da1.data_ptr(), mid, y1.data_ptr(), rows, mid, r["ss"].data_ptr(), r["mr"].data_ptr(), 2,
                  1.0, part.data_ptr(), part.numel(), self.bufs["bn.coef"].data_ptr(), dy1.data_ptr(),
                  flat.g32(bn.weight).data_ptr(), flat.g32(bn.bias).data_ptr())
        self.gemm_tr(dy1.data_ptr(), mid, 1, P.data_ptr(), ch, 1, mid, ch, rows,
                     self._epi(flat.g32(att.conv1.weight).view(mid, ch), ch, accumulate=1), self._splits(mid, ch, rows))
        self.call("pose_colsum_bf16", dy1.data_ptr(), rows, mid, mid, flat.g32(att.conv1.bias).data_ptr())
        dP = self.buf(name + ".dP", rows, ch)
        self.gemm_tr(dy1.data_ptr(), mid, 0, flat.w16(att.conv1.weight).data_ptr(), ch, 1, rows, ch, mid, self._epi(dP, ch))
        self.call("pose_coord_bwd_apply_bf16", dout.data_ptr(), G.data_ptr(), dP.data_ptr(), Bn, H, W, ch, dx.data_ptr())
        if add is not None:
            self.call("pose_add_bf16", dx.data_ptr(), add.data_ptr(), dx.numel(), dx.data_ptr())
        return dx

    # ---- blocks ------------------------------------------------------------------------------------------
    def ir_fwd(self, name, blk, x, shape):
        cm = self.cm
        mods = list(blk.conv)
        cbas = [(i, mm) for i, mm in enumerate(mods) if isinstance(mm, cm.ConvBnAct)]
        atts = [(i, mm) for i, mm in enumerate(mods) if not isinstance(mm, cm.ConvBnAct)]
        y, s = x, shape
        if len(cbas) == 3:
            y, s = self.cba_fwd(f"{name}.conv.{cbas[0][0]}", cbas[0][1], y, s)
        gated = bool(atts) and isinstance(atts[0][1], (cm.SEBlock, cm.ECABlock))
        y, s = self.cba_fwd(f"{name}.conv.{cbas[-2][0]}", cbas[-2][1], y, s,
                            pool=f"{name}.conv.{atts[0][0]}.pool" if gated else None)
        if atts:
            y = self.att_fwd(f"{name}.conv.{atts[0][0]}", atts[0][1], y, s, pool=self._pool if gated else None)
        return self.cba_fwd(f"{name}.conv.{cbas[-1][0]}", cbas[-1][1], y, s, act=None,
                            residual=x if blk.use_residual else None, ld_res=shape[3])

    def ir_bwd(self, name, blk, dout):
        cm = self.cm
        mods = list(blk.conv)
        cbas = [(i, mm) for i, mm in enumerate(mods) if isinstance(mm, cm.ConvBnAct)]
        atts = [(i, mm) for i, mm in enumerate(mods) if not isinstance(mm, cm.ConvBnAct)]
        co = self.rec[f"{name}.conv.{cbas[-1][0]}"]["co"]
        d = self.cba_bwd(f"{name}.conv.{cbas[-1][0]}", dout, co)
        res = dout if blk.use_residual else None
        first_is_dw = len(cbas) == 2
        dw_name = f"{name}.conv.{cbas[-2][0]}"
        parts = None
        if atts:
            d = self.att_bwd(f"{name}.conv.{atts[0][0]}", d, bn_next=self.rec[dw_name])
            parts = self._emit_parts
        d = self.cba_bwd(dw_name, d, d.shape[1], dx_add=res if first_is_dw else None, dz_parts=parts,
                         bn_next=None if first_is_dw else self.rec[f"{name}.conv.{cbas[0][0]}"])
        parts = self._emit_parts
        if not first_is_dw:
            d = self.cba_bwd(f"{name}.conv.{cbas[0][0]}", d, d.shape[1], dx_add=res, dz_parts=parts)
        return d

    def dual_fwd(self, name, blk, x, shape):
        cm = self.cm
        Bn, H, W, cin = shape
        cout = blk.fusion.conv.out_channels
        dense = blk.dense_path[0].conv.out_channels
        r, s = self.cba_fwd(name + ".residual_path.0", blk.residual_path[0], x, shape)
        r, s = self.cba_fwd(name + ".residual_path.1.depthwise", blk.residual_path[1].depthwise, r, s)
        r, s = self.cba_fwd(name + ".residual_path.1.pointwise", blk.residual_path[1].pointwise, r, s)
        if isinstance(blk.shortcut, cm.ConvBnAct):
            sc, _ = self.cba_fwd(name + ".shortcut", blk.shortcut, x, shape, act=None)
        else:
            sc = x
        M = s[0] * s[1] * s[2]
        cat = self.buf(name + ".cat", M, cout + dense)
        self.cba_fwd(name + ".residual_path.2", blk.residual_path[2], r, s, act=None, out=cat, ld_out=cout + dense,
                     residual=sc, ld_res=cout)
        d, sd = self.cba_fwd(name + ".dense_path.0", blk.dense_path[0], x, shape)
        d, sd = self.cba_fwd(name + ".dense_path.1.depthwise", blk.dense_path[1].depthwise, d, sd)
        self.cba_fwd(name + ".dense_path.1.pointwise", blk.dense_path[1].pointwise, d, sd, out=cat, ld_out=cout + dense,
                     col_off=cout)
        gated = blk.attention is not None and isinstance(blk.attention, (cm.SEBlock, cm.ECABlock))
        f, sf = self.cba_fwd(name + ".fusion", blk.fusion, cat, (s[0], s[1], s[2], cout + dense),
                             pool=name + ".attention.pool" if gated else None)
        if blk.attention is not None:
            f = self.att_fwd(name + ".attention", blk.attention, f, sf, pool=self._pool if gated else None)
        return f, sf

    def dual_bwd(self, name, blk, dout):
        cm = self.cm
        cout = blk.fusion.conv.out_channels
        dense = blk.dense_path[0].conv.out_channels
        ldc = cout + dense
        d = self.att_bwd(name + ".attention", dout) if blk.attention is not None else dout
        dcat = self.cba_bwd(name + ".fusion", d, cout)
        # dense path (columns cout..)
        g = self.cba_bwd(name + ".dense_path.1.pointwise", dcat, ldc, da_off=cout)
        g = self.cba_bwd(name + ".dense_path.1.depthwise", g, g.shape[1], bn_next=self.rec[name + ".dense_path.0"])
        parts = self._emit_parts
        identity = not isinstance(blk.shortcut, cm.ConvBnAct)
        dx = self.cba_bwd(name + ".dense_path.0", g, g.shape[1], dx_add=dcat if identity else None, ld_add=ldc, dz_parts=parts)
        # residual path (columns 0..cout)
        g = self.cba_bwd(name + ".residual_path.2", dcat, ldc)
        g = self.cba_bwd(name + ".residual_path.1.pointwise", g, g.shape[1])
        g = self.cba_bwd(name + ".residual_path.1.depthwise", g, g.shape[1], bn_next=self.rec[name + ".residual_path.0"])
        parts = self._emit_parts
        dx = self.cba_bwd(name + ".residual_path.0", g, g.shape[1], dx_add=dx, dz_parts=parts)
        if not identity:
            dx = self.cba_bwd(name + ".shortcut", dcat, ldc, dx_add=dx)
        return dx

    def wasp_fwd(self, x, shape):
        m, flat = self.model.wasp, self.flat
        Bn, H, W, ch = shape
        HW, M = H * W, Bn * H * W
        co = m.conv1x1.conv.out_channels
        nb = 1 + len(m.atrous_branches)
        br = self.buf("wasp.branches", nb, M, co)
        self.cba_fwd("wasp.conv1x1", m.conv1x1, x, shape, out=br[0], ld_out=co)
        for i, b in enumerate(m.atrous_branches):
            self.cba_fwd(f"wasp.atrous_branches.{i}", b, x, shape, out=br[1 + i], ld_out=co)
        pool = self.pool_sums("wasp.pool", x, Bn, HW, ch)
        mean16 = self.buf("wasp.mean", Bn, ch)
        self.call("pose_sums_to_bf16", pool.data_ptr(), pool.shape[1], Bn, ch, 1.0 / HW, mean16.data_ptr())
        g, _ = self.cba_fwd("wasp.global_branch.1", m.global_branch[1], mean16, (Bn, 1, 1, ch))
        comb = self.buf("wasp.comb", M, co)
        # bilinear interpolation of a 1x1 map is a broadcast (cnn.py:465-467)
        self.call("pose_wasp_mix_bf16", br.data_ptr(), nb, g.data_ptr(), flat.f32(m.weights).data_ptr(), Bn, HW, co,
                  comb.data_ptr())
        self.rec["wasp"] = dict(x=x, shape=shape, br=br, g=g, nb=nb, co=co)
        return self.cba_fwd("wasp.fusion", m.fusion, comb, (Bn, H, W, co))

    def wasp_bwd(self, dout):
        m, flat = self.model.wasp, self.flat
        r = self.rec["wasp"]
        Bn, H, W, ch = r["shape"]
        HW, M, co, nb = H * W, Bn * H * W, r["co"], r["nb"]
        dcomb = self.cba_bwd("wasp.fusion", dout, co)
        dbr = self.buf("wasp.dbranches", nb, M, co)
        dglob = self.ws("wasp.dglob", Bn * co)
        dots = self.ws("wasp.dots", 8)
        self.call("pose_wasp_mix_bwd_bf16", dcomb.data_ptr(), r["br"].data_ptr(), nb, r["g"].data_ptr(),
                  flat.f32(m.weights).data_ptr(), Bn, HW, co, dbr.data_ptr(), dglob.data_ptr(), dots.data_ptr(),
                  flat.g32(m.weights).data_ptr())
        dg16 = self.buf("wasp.dg16", Bn, co)
        self.call("pose_cast_f32_bf16", dglob.data_ptr(), dg16.data_ptr(), Bn * co)
        dmean = self.cba_bwd("wasp.global_branch.1", dg16, co)
        dx = self.cba_bwd("wasp.conv1x1", dbr[0], co)
        for i in range(len(m.atrous_branches)):
            dx = self.cba_bwd(f"wasp.atrous_branches.{i}", dbr[1 + i], co, dx_add=dx)
        out = self.buf("wasp.dx", M, ch)
        self.call("pose_gate_bwd_apply_bf16", dx.data_ptr(), None, dmean.data_ptr(), 1.0 / HW, Bn, HW, ch, None, out.data_ptr())
        return out

    # ---- whole model -----------------------------------------------------------------------------------
    def refresh_params(self):
        flat = self.flat
        flat.refresh_shadow()
        # packed copies follow the master weights: rebuilt every forward (one launch)
        self.call("pose_param_repack", self.repack_table.data_ptr(), self.n_repack, flat.master.data_ptr(),
                  self.pk32.data_ptr(), self.pk16.data_ptr())

    def forward(self, image, depth, kp, save=True):
        self._sp = _lib.stream_ptr()
        m, Bn, S = self.model, self.B, self.S
        c = m.config
        _lib.require_cuda(image, "image", torch.float32)
        _lib.require_cuda(depth, "depth", torch.float32)
        _lib.require_cuda(kp, "keypoints_2d", torch.float32)
        if tuple(image.shape) != (Bn, 3, S, S) or tuple(depth.shape) != (Bn, 1, S, S) or tuple(kp.shape) != (Bn, self.J, 2):
            raise ValueError(f"expected image [{Bn},3,{S},{S}], depth [{Bn},1,{S},{S}], keypoints [{Bn},{self.J},2]")
        self.launches = 0
        self.rec = {}
        if self.wsbuf is None:
            # first step: slots are assigned as the layers run; the exact size is known after the first backward
            self.wsbuf = torch.zeros(32 << 20, dtype=torch.float32, device=self.dev)
            self._ws_exact = False
        self.wsbuf.zero_()
        self.refresh_params()
        self.step_count += 1
        c0p = self.pk["conv1.0"][3]                # 21 input channels padded to 32 (64-byte pixels)
        x0 = self.buf("x0", Bn, S, S, c0p)
        self.call("pose_cnn_input_pack_ex", image.data_ptr(), depth.data_ptr(), kp.data_ptr(), Bn, S, self.J,
                  float(c.heatmap_sigma), c0p, x0.data_ptr())
        x, s = self.cba_fwd("conv1.0", m.conv1[0], x0, (Bn, S, S, c0p))
        x, s = self.cba_fwd("conv1.1", m.conv1[1], x, s)
        self.blocks = []
        for i, stage in enumerate(m.stages):
            for j, blk in enumerate(stage):
                name = f"stages.{i}.{j}"
                dual = isinstance(blk, self.cm.DualPathBlock)
                x, s = self.dual_fwd(name, blk, x, s) if dual else self.ir_fwd(name, blk, x, s)
                self.blocks.append((name, blk, dual))
        x, s = self.wasp_fwd(x, s)
        gp = int(c.global_pool_size)
        H = s[1]
        self.pooled = None
        if H == 2 * gp:
            y = self.buf("gf.pool", Bn * gp * gp, s[3])
            self.call("pose_avgpool2x2_bf16", x.data_ptr(), Bn, H, H, s[3], y.data_ptr())
            self.pooled = (Bn, H, H, s[3])
            x, s = y, (Bn, gp, gp, s[3])
        elif H > gp:
            y = self.buf("gf.pool", Bn * gp * gp, s[3])
            self.call("pose_adaptive_avgpool_bf16", x.data_ptr(), Bn, H, H, s[3], gp, gp, y.data_ptr())
            self.pooled = (Bn, H, H, s[3])
            x, s = y, (Bn, gp, gp, s[3])
        elif H != gp:
            raise NotImplementedError(f"AdaptiveAvgPool2d({gp}) from a smaller {H}x{H} map (up-sampling) is not built")
        x, s = self.cba_fwd("global_features.1", m.global_features[1], x, s)
        ch = s[3]
        eca = m.global_features[2]
        pool = self.pool_sums("gf.pool_sums", x, Bn, gp * gp, ch)
        gate = self.buf("gf.gate", Bn, ch, dtype=torch.float32)
        feat = self.buf("gf.feat", Bn, ch)
        k = eca.conv.weight.shape[-1]
        self.call("pose_eca_gate", pool.data_ptr(), pool.shape[1], 1.0 / (gp * gp), self.flat.f32(eca.conv.weight).data_ptr(),
                  k, Bn, ch, gate.data_ptr(), feat.data_ptr())
        self.rec["gf"] = dict(pool=pool, gate=gate, shape=s, k=k)
        # regression head: [Linear -> act -> Dropout] x n -> Linear (common.py:73-89)
        lins = [mod[0] if isinstance(mod, nn.Sequential) else mod for mod in m.pose_head.decoder]
        h = feat
        for i, lin in enumerate(lins):
            last = i == len(lins) - 1
            n_out, n_in = lin.weight.shape
            out = self.buf(f"head{i}", Bn, n_out, dtype=torch.float32 if last else torch.bfloat16)
            u = None if last else self.buf(f"head{i}.u", Bn, n_out)
            e = self._epi(out, n_out, bias=self.flat.f32(lin.bias), act=0 if last else self.act, preact=u)
            self.gemm(h.data_ptr(), n_in, self.flat.w16(lin.weight).data_ptr(), n_in, Bn, n_out, n_in, e)
            h = out
            if not last and self.drop_p > 0.0:
                hd = self.buf(f"head{i}.drop", Bn, n_out)
                self.call("pose_dropout_bf16", h.data_ptr(), h.numel(), self.drop_p, self._seed(i), hd.data_ptr())
                h = hd
        self.lins = lins
        if self.bn_tracked:
            torch._foreach_add_(self.bn_tracked, 1)
        self.saved = True
        return h

    def backward(self, dout, section_done=None):
        self._sp = _lib.stream_ptr()
        if not getattr(self, "saved", False):
            raise RuntimeError("backward needs a training-mode forward on this plan first")
        m, Bn, flat, b = self.model, self.B, self.flat, self.bufs
        if not flat.grads_attached():
            flat.grad.zero_()
            flat.attach_grads()
        _lib.require_cuda(dout, "grad_output", torch.float32)
        hi = [flat.numel]

        def done(first_param):
            if section_done is not None:
                lo = flat.index[id(first_param)] if first_param is not None else 0
                if lo < hi[0]:
                    section_done(flat, lo, hi[0])
                    hi[0] = lo
        n_out = self.J * 3
        ld = _ceil8(n_out)
        dy = self.buf("d.out", Bn, ld)
        self.call("pose_cast_f32_bf16_2d", dout.data_ptr(), n_out, Bn, n_out, dy.data_ptr(), ld)
        lins = self.lins
        for i in range(len(lins) - 1, -1, -1):
            lin = lins[i]
            n_o, n_i = lin.weight.shape
            if i > 0:
                x = b[f"head{i - 1}.drop"] if self.drop_p > 0.0 else b[f"head{i - 1}"]
                u = b[f"head{i - 1}.u"]
            else:
                x, u = b["gf.feat"], None
            self.gemm_tr(dy.data_ptr(), ld, 1, x.data_ptr(), n_i, 1, n_o, n_i, Bn, self._epi(flat.g32(lin.weight), n_i, accumulate=1))
            self.call("pose_colsum_bf16", dy.data_ptr(), Bn, n_o, ld, flat.g32(lin.bias).data_ptr())
            dx = self.buf(f"d.head{i}", Bn, n_i)
            e = self._epi(dx, n_i, act=ACT_GRAD[self.act] if u is not None else 0, residual=u, ldr=n_i)
            self.gemm_tr(dy.data_ptr(), ld, 0, flat.w16(lin.weight).data_ptr(), n_i, 1, Bn, n_i, n_o, e)
            if i > 0 and self.drop_p > 0.0:     # dropout mask of the forward (same seed); commutes with act'
                self.call("pose_dropout_bf16", dx.data_ptr(), dx.numel(), self.drop_p, self._seed(i - 1), dx.data_ptr())
            dy, ld = dx, n_i
        done(lins[0].weight)
        # ECABlock + AdaptiveAvgPool2d(1): feat = mean * gate
        g = self.rec["gf"]
        _, gp, _, ch = g["shape"]
        eca = m.global_features[2]
        dmean = self.buf("gf.dmean", Bn, ch)
        self.call("pose_eca_bwd", None, dy.data_ptr(), g["gate"].data_ptr(), g["pool"].data_ptr(), g["pool"].shape[1],
                  1.0 / (gp * gp), flat.f32(eca.conv.weight).data_ptr(), g["k"], Bn, ch, 1, dmean.data_ptr(),
                  flat.g32(eca.conv.weight).data_ptr())
        da = self.buf("gf.da", Bn * gp * gp, ch)
        self.call("pose_gate_bwd_apply_bf16", None, None, dmean.data_ptr(), 1.0 / (gp * gp), Bn, gp * gp, ch, None, da.data_ptr())
        d = self.cba_bwd("global_features.1", da, ch)
        if self.pooled is not None:
            _, H, W, cc = self.pooled
            dfull = self.buf("gf.dpool", Bn * H * W, cc)
            if H == 2 * gp:
                self.call("pose_avgpool2x2_bwd_bf16", d.data_ptr(), Bn, H, W, cc, dfull.data_ptr())
            else:
                self.call("pose_adaptive_avgpool_bwd_bf16", d.data_ptr(), Bn, H, W, cc, gp, gp, dfull.data_ptr())
            d = dfull
        done(m.global_features[1].conv.weight)
        d = self.wasp_bwd(d)
        done(m.wasp.weights)
        for name, blk, dual in reversed(self.blocks):
            d = self.dual_bwd(name, blk, d) if dual else self.ir_bwd(name, blk, d)
            done(next(blk.parameters()))
        d = self.cba_bwd("conv1.1", d, d.shape[1])
        self.cba_bwd("conv1.0", d, d.shape[1], need_dx=False)
        done(None)
        if not self._ws_exact:
            self.wsbuf = torch.zeros(self._ws_off, dtype=torch.float32, device=self.dev)
            self._ws_exact = True
