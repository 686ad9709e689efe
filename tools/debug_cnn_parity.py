"""Whole-step parity numbers of the CNN training step against the fp32 oracle, next to torch.autocast(bfloat16):
python tools/debug_cnn_parity.py [B]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pose = importlib.import_module("3dhumanposeestimation_b200")
from oracle import torch_models as tm
DEV = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
m = pose.CNNPoseEstimation(pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256, regression_dropout=0.0))
sd = tm.fill_state_dict(m.state_dict(), seed=5)
m.load_state_dict(sd)
m = m.to(DEV).train()
sd = {k: v.to(DEV) for k, v in sd.items()}
g = torch.Generator().manual_seed(21)
img, dep = torch.rand(B, 3, 256, 256, generator=g).to(DEV), torch.rand(B, 1, 256, 256, generator=g).to(DEV)
kp = (torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05).to(DEV)
gt = (torch.randn(B, 17, 3, generator=g) * 300).to(DEV)
crit = pose.ComprehensivePoseLoss()
pred = m(img, dep, kp)
total, _ = crit(pred, gt)
total.backward()
names = [n for n, _ in m.named_parameters()]
iu = torch.triu_indices(17, 17, 1, device=DEV)
pd = lambda t: torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]

def oracle_step(autocast):
    sdg = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        po, stats = tm.cnn_forward(sdg, m.config, img, dep, kp, train=True, return_stats=True)
    po = po.float()
    d = po - gt
    (((d ** 2).mean() + d.abs().mean() + 100.0 * (pd(po) - pd(gt)).abs().mean() + d[:, 0].abs().mean())).backward()
    return po.detach(), {n: sdg[n].grad.double() for n in names}

torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
p32, g32 = oracle_step(False)
p16, g16 = oracle_step(True)
print(f"B={B} MPJPE vs fp32: ours {pose.utils.compute_mpjpe(pred.detach(), p32).item():.3f} mm, autocast {pose.utils.compute_mpjpe(p16, p32).item():.3f} mm; "
      f"|pred| ~ {p32.abs().mean().item():.1f}")
ours, auto, nr = [], [], []
gmax = max(v.norm().item() for v in g32.values())
for n, p in m.named_parameters():
    r = g32[n]; rn = r.norm().item() + 1e-12
    if rn < 1e-5 * gmax:
        continue          # analytically-zero gradients (conv bias in front of a BatchNorm) hold rounding noise only
    ours.append(((p.grad.double() - r).norm().item() / rn, n)); auto.append(((g16[n] - r).norm().item() / rn, n))
    nr.append((abs(p.grad.double().norm().item() - rn) / rn, n))
big = [(o, n) for o, n in ours if g32[n].norm().item() > 1e-6 * max(v.norm().item() for v in g32.values())]
print(f"gradient error (relative, per parameter): ours mean {sum(o for o, _ in ours) / len(ours):.4f} max {max(ours)[0]:.4f} ({max(ours)[1]}); "
      f"autocast mean {sum(o for o, _ in auto) / len(auto):.4f} max {max(auto)[0]:.4f} ({max(auto)[1]})")
print(f"gradient norm deviation: mean {sum(o for o, _ in nr) / len(nr):.4f} max {max(nr)[0]:.4f} ({max(nr)[1]})")
nra = []
for n, p in m.named_parameters():
    rn = g32[n].norm().item()
    if rn >= 1e-5 * gmax:
        nra.append((abs(g16[n].norm().item() - rn) / rn, n))
print(f"autocast gradient norm deviation: mean {sum(o for o, _ in nra) / len(nra):.4f} max {max(nra)[0]:.4f} ({max(nra)[1]})")
print("worst 6 norm dev ours:", [(round(o, 3), n) for o, n in sorted(nr, reverse=True)[:6]])
print("worst 6 ours:", [(round(o, 3), n) for o, n in sorted(ours, reverse=True)[:6]])
print("worst 6 auto:", [(round(o, 3), n) for o, n in sorted(auto, reverse=True)[:6]])
