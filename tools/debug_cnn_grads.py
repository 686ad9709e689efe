import importlib, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
from oracle import torch_models as tm
DEV = "cuda"
B = int(os.environ.get("B", "4"))
structured = os.environ.get("STRUCT", "1") == "1"
cfg = pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256, regression_dropout=0.0)
m = pose.CNNPoseEstimation(cfg)
sd = tm.fill_state_dict(m.state_dict(), seed=5); m.load_state_dict(sd); m = m.to(DEV).train()
sd = {k: v.to(DEV) for k, v in sd.items()}
g = torch.Generator().manual_seed(21)
img = torch.rand(B, 3, 256, 256, generator=g); dep = torch.rand(B, 1, 256, 256, generator=g)
if structured:
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 256), torch.linspace(0, 1, 256), indexing="ij")
    for b in range(B):
        a = torch.rand(6, generator=g)
        pat = 0.5 + 0.5 * torch.sin(6.28 * (a[0] * 3 * xx + a[1] * 3 * yy) + 6.28 * a[2])
        img[b] = (0.25 * img[b] + 0.75 * pat * a[3:6, None, None]).clamp(0, 1)
        dep[b] = (0.2 * dep[b] + 0.8 * (a[4] * xx + (1 - a[4]) * yy) * a[5]).clamp(0, 1)
img, dep = img.to(DEV), dep.to(DEV)
kp = (torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05).to(DEV); gt = (torch.randn(B, 17, 3, generator=g) * 300).to(DEV)
crit = pose.ComprehensivePoseLoss()
pred = m(img, dep, kp); total, _ = crit(pred, gt); total.backward()
names = [n for n, _ in m.named_parameters()]
sdg = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
po = tm.cnn_forward(sdg, cfg, img, dep, kp, train=True)
print("MPJPE product vs fp32 oracle: %.4f mm   |out| max %.1f" % (pose.utils.compute_mpjpe(pred.detach(), po.detach()).item(), po.abs().max().item()))
d = po - gt; iu = torch.triu_indices(17, 17, 1, device=DEV)
pd = lambda t: torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]
lo = (d ** 2).mean() + d.abs().mean() + 100.0 * (pd(po) - pd(gt)).abs().mean() + d[:, 0].abs().mean()
lo.backward()
print("loss", total.item(), lo.item())
gmax = max(sdg[n].grad.norm().item() for n in names)
rows = []
for n, p in m.named_parameters():
    a, r = p.grad.double(), sdg[n].grad.double()
    rows.append(((a - r).norm().item() / (r.norm().item() + 1e-4 * gmax), n, r.norm().item(), a.norm().item()))
for r in rows[::-1]:
    if r[0] > float(os.environ.get("THR", "0.1")): print("%.3f %-62s ref %.3e ours %.3e" % r)
print("mean rel", sum(r[0] for r in rows) / len(rows), "max", max(rows)[:2])
# ---- plain PyTorch bf16 autocast of the same oracle, for scale
sd16 = {k: (v.clone().requires_grad_() if k in names else v.clone()) for k, v in sd.items()}
with torch.autocast("cuda", dtype=torch.bfloat16):
    p16 = tm.cnn_forward(sd16, cfg, img, dep, kp, train=True)
p16 = p16.float()
print("MPJPE torch-autocast-bf16 vs fp32 oracle: %.4f mm" % pose.utils.compute_mpjpe(p16.detach(), po.detach()).item())
d = p16 - gt
l16 = (d ** 2).mean() + d.abs().mean() + 100.0 * (pd(p16) - pd(gt)).abs().mean() + d[:, 0].abs().mean()
l16.backward()
r16 = []
for n in names:
    a, r = sd16[n].grad.double(), sdg[n].grad.double()
    r16.append((a - r).norm().item() / (r.norm().item() + 1e-4 * gmax))
print("autocast grads: mean rel", sum(r16) / len(r16), "max", max(r16))
