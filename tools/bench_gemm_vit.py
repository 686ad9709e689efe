"""ViT-shaped GEMMs on the tcgen05 kernel: separates main-loop cost from epilogue cost."""
import importlib, os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
lib = pose._lib.lib(); dev = torch.device("cuda", 0)
sp = torch.cuda.current_stream().cuda_stream

def timeit(f, reps=10):
    for _ in range(3): assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def fwd(M, N, K, act, res=False, preact=False):
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16(); b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); r = torch.randn(M, N, device=dev).bfloat16(); u = torch.empty_like(out)
    e = pose._lib.PoseGemmEpilogue(); e.bias = b.data_ptr(); e.C = out.data_ptr(); e.ldc = N; e.act = act; e.out_dtype = 1; e.out_scale = 1.0
    if res: e.residual, e.ldr, e.res_scale = r.data_ptr(), N, 1.0
    if preact: e.preact = u.data_ptr()
    us = timeit(lambda: lib.pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e), sp))
    print(f"fwd   M={M} N={N} K={K} act={act} res={int(res)} preact={int(preact)}: {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s")

def dgrad(M, N, K, act=0):   # dX[M,K'] = dY[M,N'] W[N',K']  -> gemm M, N=K', K=N'
    dy = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(K, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); u = torch.randn(M, N, device=dev).bfloat16()
    e = pose._lib.PoseGemmEpilogue(); e.C = out.data_ptr(); e.ldc = N; e.act = act; e.out_dtype = 1; e.out_scale = 1.0
    if act: e.residual, e.ldr = u.data_ptr(), N
    us = timeit(lambda: lib.pose_gemm_bf16_tr(dy.data_ptr(), K, 0, w.data_ptr(), N, 1, M, N, K, 1, C.byref(e), sp))
    print(f"dgrad M={M} N={N} K={K} act={act}: {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s")

def wgrad(M, N, K, splits):  # dW[N',K'] = dY^T X: gemm M=N', N=K', K=M
    dy = torch.randn(K, M, device=dev).bfloat16(); x = torch.randn(K, N, device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev)
    e = pose._lib.PoseGemmEpilogue(); e.C = out.data_ptr(); e.ldc = N; e.out_dtype = 0; e.out_scale = 1.0; e.accumulate = 1
    us = timeit(lambda: lib.pose_gemm_bf16_tr(dy.data_ptr(), M, 1, x.data_ptr(), N, 1, M, N, K, splits, C.byref(e), sp))
    print(f"wgrad M={M} N={N} K={K} splits={splits}: {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s")

T = 64 * 257
fwd(T, 3072, 768, 0); fwd(T, 3072, 768, 3); fwd(T, 3072, 768, 3, preact=True); fwd(T, 3072, 768, 2)
fwd(T, 768, 3072, 0); fwd(T, 768, 3072, 0, res=True); fwd(T, 2304, 768, 0); fwd(T, 768, 768, 0, res=True)
dgrad(T, 3072, 768, 0); dgrad(T, 3072, 768, 5); dgrad(T, 768, 3072); dgrad(T, 768, 2304); dgrad(T, 768, 768)
for s in (1, 2, 4, 8): wgrad(768, 3072, T, s)
for s in (1, 2, 4): wgrad(3072, 768, T, s)
for s in (1, 2, 4, 8): wgrad(2304, 768, T, s)
for s in (1, 2, 4, 8): wgrad(768, 768, T, s)
T2 = 64 * 273
for s in (1, 2, 4): wgrad(768, 3072, T2, s)
for s in (2, 4, 8): wgrad(768, 768, T2, s)
fwd(8192, 8192, 8192, 0)
