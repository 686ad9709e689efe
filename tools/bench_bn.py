"""Micro-benchmark of the training-mode BatchNorm passes on the CNN's layer shapes (rows x channels):
python tools/bench_bn.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pose = importlib.import_module("3dhumanposeestimation_b200")
lib = pose._lib.lib()
dev = torch.device("cuda", 0)
sp = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=10):
    for _ in range(2): assert fn() == 0
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                                   # the 126 MB L2 does not hold the previous repetition
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3

def run(M, C, act):
    y = torch.randn(M, C, device=dev).bfloat16()
    da = torch.randn(M, C, device=dev).bfloat16()
    res = torch.randn(M, C, device=dev).bfloat16()
    out = torch.empty_like(y)
    ss = torch.randn(2 * C, device=dev)
    mr = torch.rand(2 * C, device=dev) + 0.5
    part = torch.empty(8 << 20, device=dev)
    coef = torch.empty(2 * C, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    el = M * C
    t = timed(lambda: lib.pose_bn_apply_bf16(y.data_ptr(), M, C, ss.data_ptr(), act, 1.0, None, 0, out.data_ptr(), C, sp()))
    t2 = timed(lambda: lib.pose_bn_apply_bf16(y.data_ptr(), M, C, ss.data_ptr(), act, 1.0, res.data_ptr(), C, out.data_ptr(), C, sp()))
    t3 = timed(lambda: lib.pose_bn_bwd_bf16(da.data_ptr(), C, y.data_ptr(), M, C, ss.data_ptr(), mr.data_ptr(), act, 1.0,
                                            part.data_ptr(), part.numel(), coef.data_ptr(), out.data_ptr(), dg.data_ptr(), db.data_ptr(), sp()))
    t4 = timed(lambda: lib.pose_bn_stats_bf16(y.data_ptr(), M, C, C, part.data_ptr(), part.numel(), sp()))
    print(f"M={M:8d} C={C:5d} act={act}: apply {t:7.1f} us {4 * el / t / 1e6:5.2f} TB/s | apply+res {t2:7.1f} us {6 * el / t2 / 1e6:5.2f} TB/s | "
          f"bwd (reduce+coef+apply) {t3:7.1f} us {10 * el / t3 / 1e6:5.2f} TB/s | stats {t4:7.1f} us {2 * el / t4 / 1e6:5.2f} TB/s")

for M, C, act in [(131072, 768, 2), (32768, 3072, 2), (32768, 512, 2), (2097152, 64, 2), (524288, 384, 2), (524288, 128, 0),
                  (131072, 256, 0), (32768, 768, 0)]:
    run(M, C, act)
