"""ctypes binding of libpose_b200.so -- the only way the Python host side reaches the GPU kernels.

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads the CUDA runtime the library links against)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POSE_B200_LIB") or os.path.join(_HERE, "libpose_b200.so")   # override: A/B measurements only

c_int, c_float, c_size_t, c_void_p = C.c_int, C.c_float, C.c_size_t, C.c_void_p


class PoseBnFuse(C.Structure):
    """Mirror of `pose_bn_fuse` in include/pose_b200.h."""

    _fields_ = [("partials", C.c_void_p), ("cap_floats", C.c_int64), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("eps", C.c_float), ("momentum", C.c_float), ("count", C.c_int64), ("mean_rstd", C.c_void_p),
                ("scale_shift", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p)]


class PoseGemmEpilogue(C.Structure):
    """Mirror of `pose_gemm_epilogue` in include/pose_b200.h."""

    _fields_ = [("bias", C.c_void_p), ("residual", C.c_void_p), ("C", C.c_void_p), ("ldc", C.c_int32),
                ("ldr", C.c_int32), ("act", C.c_int32), ("out_dtype", C.c_int32), ("out_scale", C.c_float),
                ("res_scale", C.c_float), ("preact", C.c_void_p), ("accumulate", C.c_int32), ("reserved", C.c_int32),
                ("drop_seed", C.c_uint64), ("drop_p", C.c_float), ("reserved2", C.c_int32),
                ("bn", C.POINTER(PoseBnFuse))]


class PoseRepackEntry(C.Structure):
    """Mirror of `pose_repack_entry` in include/pose_b200.h."""

    _fields_ = [("src", C.c_int64), ("dst", C.c_int64), ("kind", C.c_int32), ("d0", C.c_int32), ("d1", C.c_int32),
                ("d2", C.c_int32), ("d3", C.c_int32), ("pad", C.c_int32)]


class PoseStepState(C.Structure):
    """Mirror of `pose_step_state` in include/pose_b200.h (lives in DEVICE memory; this mirror is for sizes / offsets)."""

    _fields_ = [("drop_key", C.c_uint32), ("adam_step", C.c_int32), ("counter", C.c_uint32), ("reserved", C.c_uint32)]


class PoseAugLaunch(C.Structure):
    """Mirror of `pose_aug_launch` in include/pose_b200.h."""

    _fields_ = [(n, C.c_int32) for n in ("max_out_h", "max_out_w", "max_rot_rows", "max_band_rows", "max_ksize",
                                        "smem_bytes", "cluster", "reserved")]


POSE_AUG_PLAN_BYTES = 256

# name -> (restype, argtypes); must list every symbol include/pose_b200.h declares
SIGNATURES = {
    "pose_b200_abi_version": (c_int, []),
    "pose_b200_error_string": (C.c_char_p, [c_int]),
    "pose_loss_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pose_loss_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, C.POINTER(c_float), c_void_p, c_void_p, c_float,
                                  c_void_p, c_size_t, c_void_p]),
    "pose_heatmap_patchify_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "pose_heatmap_render": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    "pose_augment_plan": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, C.POINTER(PoseAugLaunch)]),
    "pose_augment_workspace_bytes": (c_size_t, [c_int, c_int, c_int, C.POINTER(PoseAugLaunch)]),
    "pose_augment_batch": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   C.POINTER(PoseAugLaunch), c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pose_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_void_p]),
    "pose_cast_f32_bf16": (c_int, [c_void_p, c_void_p, C.c_long, c_void_p]),
    "pose_gemm_bf16_ex": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, C.POINTER(PoseGemmEpilogue),
                                  c_void_p]),
    "pose_gemm_bf16_tr": (c_int, [c_void_p, C.c_long, c_int, c_void_p, C.c_long, c_int, c_int, c_int, c_int, c_int,
                                  C.POINTER(PoseGemmEpilogue), c_void_p]),
    "pose_conv2d_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_int, C.POINTER(PoseGemmEpilogue), c_void_p]),
    "pose_cnn_input_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "pose_dwconv3x3_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                    c_void_p, c_int, c_void_p]),
    "pose_dwconv3x3_pool_parts": (c_int, [c_int, c_int, c_int]),
    "pose_pool_sum_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pose_se_gate": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                             c_void_p]),
    "pose_eca_gate": (c_int, [c_void_p, c_int, c_float, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pose_channel_affine_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, C.c_long, c_int, c_void_p, c_void_p]),
    "pose_coord_pool_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_coord_apply_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_avgpool2x2_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_adaptive_avgpool_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_sums_to_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "pose_layernorm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_float, C.c_long, c_int, C.c_long, C.c_long, C.c_long,
                                    C.c_long, c_int, c_void_p, c_void_p]),
    "pose_token_concat_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                       c_void_p]),
    "pose_patchify_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_attention_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, C.c_long,
                                    C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, C.c_long, c_float, c_void_p,
                                    c_float, C.c_uint64, c_void_p]),
    "pose_layernorm_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_float, C.c_long, c_int, C.c_long, C.c_long,
                                        C.c_long, C.c_long, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pose_colsum_bf16": (c_int, [c_void_p, C.c_long, c_int, C.c_long, c_void_p, c_void_p]),
    "pose_cast_f32_bf16_2d": (c_int, [c_void_p, C.c_long, C.c_long, c_int, c_void_p, C.c_long, c_void_p]),
    "pose_batch_rowsum_bf16": (c_int, [c_void_p, c_int, C.c_long, C.c_long, c_int, c_int, c_void_p, c_void_p]),
    "pose_token_slice_bf16": (c_int, [c_void_p, c_int, C.c_long, C.c_long, c_int, c_int, c_void_p, c_void_p]),
    "pose_attention_bwd_bf16": (c_int, [c_void_p] * 10 + [c_int] * 5 + [C.c_long] * 16 + [c_float, c_float, C.c_uint64,
                                                                                            c_void_p]),
    "pose_cnn_input_pack_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "pose_conv2d_wgrad_bf16": (c_int, [c_void_p, c_void_p] + [c_int] * 10 + [c_void_p, c_int, c_void_p]),
    "pose_bn_stats_bf16": (c_int, [c_void_p, C.c_long, c_int, C.c_long, c_void_p, C.c_long, c_void_p]),
    "pose_bn_finalize": (c_int, [c_void_p, C.c_long, C.c_long, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "pose_bn_finalize_parts": (c_int, [c_void_p, c_int, C.c_long, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "pose_dwconv3x3_bn_stats_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                             C.c_long, c_void_p]),
    "pose_bn_apply_bf16": (c_int, [c_void_p, C.c_long, c_int, c_void_p, c_int, c_float, c_void_p, C.c_long, c_void_p, C.c_long,
                                   c_void_p]),
    "pose_bn_apply_pool_bf16": (c_int, [c_void_p, c_int, C.c_long, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "pose_bn_bwd_bf16": (c_int, [c_void_p, C.c_long, c_void_p, C.c_long, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p,
                                 C.c_long, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pose_dwconv3x3_bnbwd_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                          c_void_p, C.c_long, c_void_p]),
    "pose_gate_bwd_apply_bn_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, C.c_long, c_int, c_void_p, c_void_p,
                                            c_int, c_void_p, c_void_p, C.c_long, c_void_p, c_void_p]),
    "pose_bn_bwd_from_dz_bf16": (c_int, [c_void_p, C.c_long, c_void_p, C.c_long, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pose_dwconv3x3_bwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "pose_gate_bwd_reduce_bf16": (c_int, [c_void_p, c_void_p, c_int, C.c_long, c_int, c_void_p, c_void_p]),
    "pose_gate_bwd_apply_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, C.c_long, c_int, c_void_p, c_void_p,
                                         c_void_p]),
    "pose_sigmoid_bwd": (c_int, [c_void_p, c_void_p, C.c_long, c_void_p, c_void_p]),
    "pose_eca_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_void_p, c_void_p]),
    "pose_coord_bwd_reduce_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_coord_bwd_apply_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_wasp_mix_bf16": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, C.c_long, c_int, c_void_p, c_void_p]),
    "pose_wasp_mix_bwd_bf16": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, C.c_long, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "pose_avgpool2x2_bwd_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_adaptive_avgpool_bwd_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_scatter_strided_add_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pose_add_bf16": (c_int, [c_void_p, c_void_p, C.c_long, c_void_p, c_void_p]),
    "pose_dropout_bf16": (c_int, [c_void_p, C.c_long, c_float, C.c_uint64, c_void_p, c_void_p]),
    "pose_param_repack": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pose_eval_metrics": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pose_collate_pad": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pose_infer_prep": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_float,
                                c_void_p, c_void_p, c_void_p]),
    "pose_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_long, c_float, c_float, c_float,
                                c_float, c_float, c_int, c_float, c_int, c_void_p]),
    "pose_resize_workspace_bytes": (C.c_size_t, [c_int, c_int, c_int, c_int]),
    "pose_resize_bilinear_aa": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        C.c_size_t, c_void_p, c_void_p]),
    "pose_step_state_bind": (c_int, [c_void_p]),
    "pose_step_tick": (c_int, [c_void_p, c_void_p]),
    "pose_adamw_step_g16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_long, c_float, c_float,
                                    c_float, c_float, c_float, c_int, c_float, c_int, c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python 3dhumanposeestimation_b200/build.py` "
                "(this package has no CPU or eager fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


_bound_state = None     # the device tensor (int32[4] = pose_step_state) currently bound, or None


def bind_step_state(state) -> None:
    """pose_step_state_bind: while bound, dropout keys and AdamW's step count come from device memory (CUDA-graph replay of
    the training step).  The binding is a host-side pointer read when a kernel is launched (or captured)."""
    global _bound_state
    code = lib().pose_step_state_bind(state.data_ptr() if state is not None else None)
    _bound_state = state
    check(code, "pose_step_state_bind")


def step_state_bound() -> bool:
    return _bound_state is not None


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().pose_b200_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    """The kernels only accept contiguous CUDA tensors of the exact dtype; anything else raises."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 kernels have no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t
