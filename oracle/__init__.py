"""CPU oracle for the pose hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package.  The product package never does: it fails loudly
when its CUDA library is missing instead of falling back to anything here.

``pose_oracle.c`` restates the integer / fixed-point / fp32 arithmetic of the reference's
PoseAugmentor (torchvision + Pillow), GaussianHeatmapGenerator and ComprehensivePoseLoss;
``torch_models.py`` restates the fp32 model forwards with plain ``torch.nn.functional`` calls.
Parity of the restatements is pinned by ``tests/golden/*.npz``, generated from the live
reference by ``oracle/gen_golden.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpose_oracle.so")

AUG_FLIP, AUG_ROTATE, AUG_SCALE, AUG_TRANSLATE, AUG_COLOR = 1, 2, 4, 8, 16
AUG_ALL = 31


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pose_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpose_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_mpjpe.restype = C.c_float
        _lib.oracle_pa_mpjpe.restype = C.c_float
        _lib.oracle_rotate_matrix.restype = C.c_int
        _lib.oracle_grey_mean_u8.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _u8(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


# ---- stage-level entry points (HWC uint8 images) ------------------------------------
def quantize_u8(x) -> np.ndarray:
    x = _f32(x)
    out = np.empty(x.shape, np.uint8)
    lib().oracle_quantize_u8(_p(x), C.c_long(x.size), _p(out))
    return out


def rotate_matrix(angle_deg: float, W: int, H: int):
    a = np.zeros(6, np.float64)
    mode = lib().oracle_rotate_matrix(C.c_double(angle_deg), C.c_int(W), C.c_int(H), _p(a))
    return mode, a


def _hwc(img):
    img = _u8(img)
    if img.ndim == 2:
        return img, img.shape[0], img.shape[1], 1
    return img, img.shape[0], img.shape[1], img.shape[2]


def affine_bilinear(img, a) -> np.ndarray:
    img, H, W, Cc = _hwc(img)
    a = np.ascontiguousarray(a, np.float64)
    out = np.empty_like(img)
    lib().oracle_affine_bilinear_u8(_p(img), H, W, Cc, _p(a), _p(out))
    return out


def affine_nearest_fixed(img, a) -> np.ndarray:
    img, H, W, Cc = _hwc(img)
    a = np.ascontiguousarray(a, np.float64)
    out = np.empty_like(img)
    lib().oracle_affine_nearest_fixed_u8(_p(img), H, W, Cc, _p(a), _p(out))
    return out


def scale_affine_nearest(img, out_hw, a) -> np.ndarray:
    img, H, W, Cc = _hwc(img)
    a = np.ascontiguousarray(a, np.float64)
    oh, ow = out_hw
    out = np.empty((oh, ow) + img.shape[2:], np.uint8)
    lib().oracle_scale_affine_nearest_u8(_p(img), H, W, Cc, oh, ow, _p(a), _p(out))
    return out


def resize_bilinear_aa(img, out_hw) -> np.ndarray:
    img, H, W, Cc = _hwc(img)
    oh, ow = out_hw
    out = np.empty((oh, ow) + img.shape[2:], np.uint8)
    lib().oracle_resize_bilinear_aa_u8(_p(img), H, W, Cc, oh, ow, _p(out))
    return out


def brightness(img, factor: float) -> np.ndarray:
    out = _u8(img).copy()
    lib().oracle_brightness_u8(_p(out), C.c_long(out.size), C.c_double(factor))
    return out


def contrast(img, factor: float) -> np.ndarray:
    out = _u8(img).copy()
    assert out.ndim == 3 and out.shape[2] == 3
    lib().oracle_contrast_u8(_p(out), C.c_long(out.shape[0] * out.shape[1]), C.c_double(factor))
    return out


# ---- whole-sample augmentation ------------------------------------------------------
def augment_out_size(H: int, W: int, params, flags: int = AUG_ALL):
    p = np.ascontiguousarray(params, np.float64)
    oh, ow = C.c_int(), C.c_int()
    lib().oracle_augment_out_size(H, W, _p(p), flags, C.byref(oh), C.byref(ow))
    return oh.value, ow.value


def augment_sample(image, depth, kp, joints, cam, params, flags: int = AUG_ALL):
    """image [3,H,W] fp32, depth [1,H,W] or [H,W] fp32, kp [J,2], joints [J,3], cam (fx,fy,cx,cy),
    params = (flip, angle_deg, scale, tx_frac, ty_frac, brightness, contrast[, _]).
    Returns dict(image [3,S,S], depth [1,S,S], keypoints_2d, joints_3d, cam)."""
    image, depth, kp, joints = _f32(image), _f32(depth), _f32(kp), _f32(joints)
    H, W = image.shape[1:]
    J = kp.shape[0]
    p = np.zeros(8, np.float64)
    p[: len(params)] = params
    cam = np.ascontiguousarray(cam, np.float64)
    oh, ow = augment_out_size(H, W, p, flags)
    io = np.empty((3, oh, ow), np.float32)
    do = np.empty((1, oh, ow), np.float32)
    ko = np.empty((J, 2), np.float32)
    jo = np.empty((J, 3), np.float32)
    co = np.empty(4, np.float64)
    lib().oracle_augment_sample(_p(image), _p(depth), _p(kp), _p(joints), _p(cam), H, W, J, _p(p), flags,
                                _p(io), _p(do), _p(ko), _p(jo), _p(co))
    return {"image": io, "depth": do, "keypoints_2d": ko, "joints_3d": jo, "cam": co}


# ---- heatmap / loss / metric --------------------------------------------------------
def heatmap(kp, hs: int, sigma: float) -> np.ndarray:
    kp = _f32(kp)
    B, J = kp.shape[:2]
    out = np.empty((B, J, hs, hs), np.float32)
    lib().oracle_heatmap(_p(kp), B, J, hs, C.c_float(sigma), _p(out))
    return out


def tensor_resize_aa(img, out_h: int, out_w: int, mode: int = 2) -> np.ndarray:
    """transforms.Resize on a float32 [C, H, W] frame (F.interpolate bilinear, antialias=True) as ATen's CPU kernel
    computes it; mode 2 = the shipped tap order (bit-equal to torch 2.11, tests/golden/resize.npz), 0 / 1 = plain / fused
    accumulation everywhere."""
    img = _f32(img)
    c, h, w = img.shape
    out = np.empty((c, out_h, out_w), np.float32)
    lib().oracle_resize_bilinear_aa(_p(img), c, h, w, out_h, out_w, mode, _p(out))
    return out


def heatmap_peak(kp, hs: int) -> np.ndarray:
    kp = _f32(kp)
    B, J = kp.shape[:2]
    out = np.empty((B, J), np.int32)
    lib().oracle_heatmap_peak(_p(kp), B, J, hs, _p(out))
    return out


def pose_loss(pred, gt, weights=(1.0, 1.0, 100.0, 1.0), want_grad: bool = True):
    """weights = (mse, l1, inter_joint, abs_root). Returns (out5, grad|None)."""
    pred, gt = _f32(pred), _f32(gt)
    B, J = pred.shape[:2]
    w = np.asarray(weights, np.float32)
    out5 = np.empty(5, np.float32)
    grad = np.empty_like(pred) if want_grad else None
    lib().oracle_pose_loss(_p(pred), _p(gt), B, J, _p(w), _p(out5), _p(grad) if want_grad else None)
    return out5, grad


def mpjpe(pred, gt) -> float:
    pred, gt = _f32(pred), _f32(gt)
    return float(lib().oracle_mpjpe(_p(pred), _p(gt), pred.shape[0], pred.shape[1]))


def pa_mpjpe(pred, gt, per_sample: bool = False):
    """utils.py:72-165 compute_pa_mpjpe (the reference's rotation convention included)."""
    pred, gt = _f32(pred), _f32(gt)
    out = np.empty(pred.shape[0], np.float32)
    m = float(lib().oracle_pa_mpjpe(_p(pred), _p(gt), pred.shape[0], pred.shape[1], _p(out)))
    return (m, out) if per_sample else m
