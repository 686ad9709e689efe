"""Large plain-epilogue GEMMs (CTA-pair path): python tools/bench_gemm_big.py"""
import importlib, os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
lib = pose._lib.lib(); dev = torch.device("cuda", 0)
sp = torch.cuda.current_stream().cuda_stream
for (M, N, K) in ((8192, 8192, 8192), (16448, 768, 3072), (16448, 3072, 768)):
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    e = pose._lib.PoseGemmEpilogue(); e.C = out.data_ptr(); e.ldc = N; e.act = 0; e.out_dtype = 1; e.out_scale = 1.0
    f = lambda: lib.pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e), sp)
    for _ in range(3): assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"M={M} N={N} K={K}: {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TFLOP/s")
