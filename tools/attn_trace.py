import importlib, sys, math, torch
sys.path.insert(0, "/root/repo")
pose = importlib.import_module("3dhumanposeestimation_b200")
lib, sp = pose._lib.lib(), pose._lib.stream_ptr
B, heads, hd, N = 64, 12, 64, 257
E_ = heads * hd
qkv = torch.randn(B, N, 3 * E_, device="cuda").bfloat16()
o = torch.empty(B, N, E_, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, heads, N, device="cuda")
p = qkv.data_ptr()
for _ in range(2):
    rc = lib.pose_attention_bf16(p, p + 2 * E_, p + 4 * E_, o.data_ptr(), B, heads, N, N, hd, 3 * E_, 3 * E_, 3 * E_, E_, N * 3 * E_, N * 3 * E_, N * 3 * E_, N * E_, 1 / math.sqrt(hd), lse.data_ptr(), 0.0, 0, sp())
    torch.cuda.synchronize()
print("rc", rc)
