"""Times the tcgen05 attention forward / backward at several sequence lengths (B=64): python tools/bench_attn.py"""
import importlib, sys, math, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
lib, sp = pose._lib.lib(), pose._lib.stream_ptr

def timed(fn, reps=10):
    for _ in range(2): assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

DROP = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
for (B, heads, hd, N) in ((64, 12, 64, 257), (64, 12, 64, 256), (64, 16, 48, 273), (64, 16, 48, 256), (64, 12, 64, 320)):
    E_ = heads * hd
    T = max(N, 273)
    qkv = torch.randn(B, T, 3 * E_, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    o = torch.empty(B, T, E_, device="cuda", dtype=torch.bfloat16); do = torch.randn_like(o)
    lse = torch.empty(B, heads, N, device="cuda"); dws = torch.empty(B, heads, N, device="cuda")
    p, dp = qkv.data_ptr(), dqkv.data_ptr()
    ld, bs, bso = 3 * E_, T * 3 * E_, T * E_
    sc = 1 / math.sqrt(hd)
    f = timed(lambda: lib.pose_attention_bf16(p, p + 2 * E_, p + 4 * E_, o.data_ptr(), B, heads, N, N, hd, ld, ld, ld, E_, bs, bs, bs, bso,
                                              sc, lse.data_ptr(), DROP, 5, sp()))
    b = timed(lambda: lib.pose_attention_bwd_bf16(p, p + 2 * E_, p + 4 * E_, o.data_ptr(), do.data_ptr(), lse.data_ptr(), dp, dp + 2 * E_,
                                                  dp + 4 * E_, dws.data_ptr(), B, heads, N, N, hd, ld, ld, ld, E_, E_, ld, ld, ld, bs, bs, bs,
                                                  bso, bso, bs, bs, bs, sc, DROP, 5, sp()))
    fl = 4.0 * B * heads * N * N * hd
    print(f"drop={DROP} B={B} heads={heads} hd={hd} N={N}: fwd {f:7.1f} us ({fl / f / 1e6:6.1f} TFLOP/s)  bwd {b:7.1f} us ({2.5 * fl / b / 1e6:6.1f} TFLOP/s)")
