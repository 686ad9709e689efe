"""Micro-benchmark of the tcgen05 GEMM: python tools/bench_gemm.py"""
import importlib, os, sys, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pose = importlib.import_module("3dhumanposeestimation_b200")
lib = pose._lib.lib()
dev = torch.device("cuda", 0)

def run(M, N, K, act=2, bias=True, out_bf16=True, reps=20):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    e = pose._lib.PoseGemmEpilogue()
    e.bias = b.data_ptr() if bias else None
    e.residual = None
    e.C = out.data_ptr(); e.ldc = N; e.ldr = 0; e.act = act; e.out_dtype = 1 if out_bf16 else 0
    e.out_scale = 1.0; e.res_scale = 0.0
    sp = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.pose_gemm_bf16_ex(a.data_ptr(), K, w.data_ptr(), K, M, N, K, C.byref(e), sp)
    for _ in range(3): assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    flops = 2.0 * M * N * K
    byt = (M * K + N * K) * 2 + M * N * (2 if out_bf16 else 4)
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    print(f"M={M:7d} N={N:5d} K={K:5d} act={act} bias={int(bias)} bf16out={int(out_bf16)}: {us:9.1f} us  {flops/us/1e6:8.1f} TFLOP/s  {byt/us/1e3:8.1f} GB/s  "
          f"{us*148/tiles:6.2f} us/tile/SM")

CONFIGS = [(524288, 384, 128, 2, True, True), (524288, 384, 128, 0, False, True), (524288, 384, 128, 0, False, False),
           (524288, 128, 64, 2, True, True), (524288, 128, 64, 0, False, True), (131072, 768, 256, 2, True, True),
           (32768, 3072, 512, 2, True, True), (32768, 512, 3072, 0, True, True), (8192, 8192, 8192, 0, False, True),
           (4096, 4096, 4096, 0, False, True), (256, 1024, 1024, 2, True, True)]
sel = [int(a) for a in sys.argv[1:]] or range(len(CONFIGS))
for i in sel:
    M, N, K, act, bias, ob = CONFIGS[i]
    run(M, N, K, act, bias, ob, reps=3 if len(sys.argv) > 1 else 20)
