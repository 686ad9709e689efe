"""Micro-benchmark of the tcgen05 GEMM on the CNN's 1x1-convolution shapes (memory-bound regime, plain bf16 epilogue,
optionally with the fused BatchNorm statistics): python tools/bench_gemm_cnn.py [bn]"""
import importlib, os, sys, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pose = importlib.import_module("3dhumanposeestimation_b200")
lib = pose._lib.lib()
dev = torch.device("cuda", 0)
WITH_BN = "bn" in sys.argv[1:]

def run(M, N, K, reps=20):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    e = pose._lib.PoseGemmEpilogue()
    e.C = out.data_ptr(); e.ldc = N; e.act = 0; e.out_dtype = 1; e.out_scale = 1.0
    keep = None
    if WITH_BN:
        f = pose._lib.PoseBnFuse()
        part = torch.empty(4 * 148 * 2 * N, device=dev); mr = torch.empty(2 * N, device=dev); ss = torch.empty(2 * N, device=dev)
        g = torch.ones(N, device=dev); b = torch.zeros(N, device=dev); rm = torch.zeros(N, device=dev); rv = torch.ones(N, device=dev)
        f.partials, f.cap_floats, f.gamma, f.beta, f.eps, f.momentum, f.count = part.data_ptr(), part.numel(), g.data_ptr(), b.data_ptr(), 1e-5, 0.1, M
        f.mean_rstd, f.scale_shift, f.running_mean, f.running_var = mr.data_ptr(), ss.data_ptr(), rm.data_ptr(), rv.data_ptr()
        e.bn = C.pointer(f); keep = (f, part, mr, ss, g, b, rm, rv)
    sp = torch.cuda.current_stream().cuda_stream
    fn = lambda: lib.pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e), sp)
    for _ in range(3): assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    byt = (M * K + N * K) * 2 + M * N * 2
    print(f"M={M:7d} N={N:5d} K={K:5d} bn={int(WITH_BN)}: {us:8.1f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s  {byt/us/1e3:7.1f} GB/s ({byt/us/1e3/6549.8:.2f} of HBM peak)")

for M, N, K in [(131072, 768, 256), (131072, 256, 768), (524288, 128, 128), (524288, 384, 128), (524288, 128, 384), (32768, 3072, 512), (32768, 512, 3072),
                (32768, 512, 512), (32768, 256, 512), (32768, 768, 512)]:
    run(M, N, K)
