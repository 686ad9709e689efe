"""Attention forward at the training batch size (several CTAs per SM, many waves), checked against PyTorch every repetition."""
import importlib, sys, math, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
lib, sp = pose._lib.lib(), pose._lib.stream_ptr
for (B, heads, hd, N, p) in ((64, 12, 64, 257, 0.0), (64, 16, 48, 273, 0.0), (64, 12, 64, 257, 0.1)):
    E_ = heads * hd
    qkv = torch.randn(B, N, 3 * E_, device="cuda").bfloat16()
    o = torch.empty(B, N, E_, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, heads, N, device="cuda")
    ptr = qkv.data_ptr()
    q, k, v = (qkv[..., i * E_:(i + 1) * E_].float().view(B, N, heads, hd).transpose(1, 2) for i in range(3))
    ref = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), -1) @ v
    ref = ref.transpose(1, 2).reshape(B, N, E_)
    for rep in range(30):
        rc = lib.pose_attention_bf16(ptr, ptr + 2 * E_, ptr + 4 * E_, o.data_ptr(), B, heads, N, N, hd, 3 * E_, 3 * E_, 3 * E_, E_,
                                     N * 3 * E_, N * 3 * E_, N * 3 * E_, N * E_, 1 / math.sqrt(hd), lse.data_ptr(), p, 1234, sp())
        assert rc == 0, rc
    torch.cuda.synchronize()
    if p == 0.0:
        print(B, heads, hd, N, "max err", (o.float() - ref).abs().max().item())
    else:
        print(B, heads, hd, N, "dropout run ok", o.float().abs().mean().item())
