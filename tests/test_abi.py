"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol the header declares."""
import ctypes
import importlib
import os
import re
import subprocess

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pose_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pose_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(pose):
    names = _declared_symbols()
    assert "pose_loss_fwd_bwd" in names and "pose_augment_batch" in names and len(names) >= 8
    lib = pose._lib.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pose_b200.h but not exported"
    # and the Python binding table covers exactly the header
    assert sorted(pose._lib.SIGNATURES) == names


def test_only_the_c_abi_is_exported(pose):
    out = subprocess.check_output(["nm", "-D", "--defined-only", pose._lib.LIB_PATH], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    extra = [s for s in syms if not s.startswith("pose_") and not s.startswith("_")]
    assert not extra, extra


def test_error_strings_and_argument_validation(pose):
    lib = pose._lib.lib()
    assert lib.pose_b200_abi_version() >= 1
    assert lib.pose_b200_error_string(0) == b"ok"
    assert b"workspace" in lib.pose_b200_error_string(-3)
    # argument errors are reported before anything touches the GPU
    w = (ctypes.c_float * 4)(1, 1, 100, 1)
    assert lib.pose_loss_fwd_bwd(None, None, 4, 17, w, None, None, 1.0, None, 0, None) == -1
    assert lib.pose_heatmap_render(None, 1, 17, 64, 2.0, None, 0, 0, 0, 0, None) == -1
    assert lib.pose_loss_workspace_bytes(256, 17) >= 16


def test_augment_plan_is_host_only(pose):
    import numpy as np
    lib = pose._lib.lib()
    B = 5
    params = np.zeros((B, 8))
    params[:, 1] = [0.0, 12.5, -29.0, 180.0, 90.0]
    params[:, 2] = [0.8, 1.0, 1.2, 0.93, 1.13]
    plan = np.zeros(B * pose._lib.POSE_AUG_PLAN_BYTES, np.uint8)
    launch = pose._lib.PoseAugLaunch()
    assert lib.pose_augment_plan(params.ctypes.data, B, 256, 256, 31, plan.ctypes.data, ctypes.byref(launch)) == 0
    assert launch.max_out_h == int(256 * 1.2) and launch.max_out_w == int(256 * 1.2)
    assert launch.cluster == 8 and 0 < launch.smem_bytes <= 227 * 1024
    assert launch.max_ksize == 5  # support = 1.25 for the 0.8x sample -> 2*ceil(1.25)+1 taps
    assert lib.pose_augment_workspace_bytes(B, 256, 256, ctypes.byref(launch)) > B * 256 * 256 * 4
    # scale < 0.25 needs more taps than the kernel keeps
    params[:, 2] = 0.2
    assert lib.pose_augment_plan(params.ctypes.data, B, 256, 256, 31, plan.ctypes.data, ctypes.byref(launch)) == -4


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "3dhumanposeestimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "pose_oracle" not in text, f


def test_cpu_tensors_raise(pose):
    import pytest
    import torch
    crit = pose.ComprehensivePoseLoss()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(torch.zeros(2, 17, 3), torch.zeros(2, 17, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pose.GaussianHeatmapGenerator(17, 64, 2.0)(torch.rand(2, 17, 2))
