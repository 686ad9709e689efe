import importlib, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
from oracle import torch_models as tm
DEV = "cuda"
gd = np.load("tests/golden/vit_256.npz")
g = torch.Generator().manual_seed(int(gd["image_seed"]))
img = torch.rand(2, 3, 256, 256, generator=g).to(DEV); dep = torch.rand(2, 1, 256, 256, generator=g).to(DEV)
kp = torch.from_numpy(gd["kp"]).to(DEV); gt = torch.from_numpy(gd["gt"]).to(DEV)
cfg = pose.ModelConfig("transformer", image_size=(256, 256), vit_pretrained=False, transformer_dropout_rate=0.0,
                       transformer_attention_dropout_rate=0.0, regression_dropout=0.0)
m = pose.TransformerPoseEstimation(cfg)
sd = tm.fill_vit_state_dict(m.state_dict(), seed=7); m.load_state_dict(sd); m = m.to(DEV).train()
sd = {k: v.to(DEV) for k, v in sd.items()}
crit = pose.ComprehensivePoseLoss()
total, _ = crit(m(img, dep, kp), gt); total.backward()
sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() and "grid" not in k else v) for k, v in sd.items()}
po = tm.vit_forward(sdg, cfg, img, dep, kp)
d = po - gt; iu = torch.triu_indices(17, 17, 1, device=DEV)
pd = lambda t: torch.linalg.norm(t[:, :, None] - t[:, None], dim=-1)[:, iu[0], iu[1]]
lo = (d ** 2).mean() + d.abs().mean() + 100.0 * (pd(po) - pd(gt)).abs().mean() + d[:, 0].abs().mean()
lo.backward()
rows = []
for n, p in m.named_parameters():
    a, r = p.grad.double(), sdg[n].grad.double()
    rows.append(((a - r).norm().item() / (r.norm().item() + 1e-12), n, r.norm().item(), a.norm().item()))
rows.sort(reverse=True)
for r in rows[:25]: print("%.4f %-70s ref %.4e ours %.4e" % r)
print("mean rel", sum(r[0] for r in rows) / len(rows))
