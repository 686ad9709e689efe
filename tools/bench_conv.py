"""Micro-benchmark of the implicit-GEMM convolution (forward / weight gradient) on the CNN's spatial-conv shapes:
python tools/bench_conv.py"""
import importlib, os, sys, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pose = importlib.import_module("3dhumanposeestimation_b200")
lib = pose._lib.lib()
dev = torch.device("cuda", 0)
sp = lambda: torch.cuda.current_stream().cuda_stream

def timed(fn, reps=10):
    for _ in range(2): assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def run(B, H, Cin, Cout, k, stride, dil, splits_list):
    pad = (k - 1) // 2 * dil
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    x = torch.randn(B, H, H, Cin, device=dev).bfloat16()
    w = torch.randn(Cout, k, k, Cin, device=dev).bfloat16()
    y = torch.empty(B, Ho, Ho, Cout, device=dev, dtype=torch.bfloat16)
    dy = torch.randn(B, Ho, Ho, Cout, device=dev).bfloat16()
    e = pose._lib.PoseGemmEpilogue()
    e.C = y.data_ptr(); e.ldc = Cout; e.act = 0; e.out_dtype = 1; e.out_scale = 1.0
    us = timed(lambda: lib.pose_conv2d_bf16(x.data_ptr(), B, H, H, Cin, w.data_ptr(), Cout, k, k, stride, dil, pad, C.byref(e), sp()))
    fl = 2.0 * B * Ho * Ho * Cout * k * k * Cin
    print(f"fwd   B={B} H={H} {Cin}->{Cout} k{k} s{stride} d{dil}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s (nominal)")
    dwk = torch.zeros(Cout, k, k, Cin, device=dev)
    for s in splits_list:
        us = timed(lambda: lib.pose_conv2d_wgrad_bf16(dy.data_ptr(), x.data_ptr(), B, H, H, Cin, Cout, k, k, stride, dil, pad, dwk.data_ptr(), s, sp()))
        print(f"wgrad splits={s:3d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s (nominal)")

run(128, 256, 32, 64, 5, 2, 1, [21, 29, 43, 84])
run(128, 256, 64, 64, 5, 2, 1, [11, 23, 46])
run(128, 128, 64, 64, 3, 1, 1, [29, 59, 60, 74, 118])
run(128, 16, 512, 512, 3, 1, 6, [3, 4, 8])
