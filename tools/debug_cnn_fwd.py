import importlib, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
from oracle import torch_models as tm
DEV = "cuda"; B = int(os.environ.get("B", "4"))
cfg = pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256, regression_dropout=0.0)
m = pose.CNNPoseEstimation(cfg)
sd = tm.fill_state_dict(m.state_dict(), seed=5); m.load_state_dict(sd); m = m.to(DEV).train()
sd = {k: v.to(DEV) for k, v in sd.items()}
g = torch.Generator().manual_seed(21)
img = torch.rand(B, 3, 256, 256, generator=g).to(DEV); dep = torch.rand(B, 1, 256, 256, generator=g).to(DEV)
kp = (torch.rand(B, 17, 2, generator=g) * 0.9 + 0.05).to(DEV)
plan = m.plan(B, img.device)
out = plan.forward(img, dep, kp)
trace = {}
with torch.no_grad():
    po = tm.cnn_forward(sd, cfg, img, dep, kp, train=True, trace=trace)
for name, ref in trace.items():
    r = plan.rec.get(name)
    if r is None: continue
    Bn, Ho, Wo, co = r["oshape"]
    a = plan.bufs.get(name + ".a")
    y = r["y"].float().view(Bn, Ho, Wo, co).permute(0, 3, 1, 2)
    msg = ""
    if a is not None:
        a = a.float().view(Bn, Ho, Wo, co).permute(0, 3, 1, 2)
        msg = "a rel %.4f" % ((a - ref).norm() / ref.norm()).item()
    mean = y.mean((0, 2, 3)); std = y.std((0, 2, 3))
    print("%-48s %s   |mean|/std of y: med %.2f max %.2f" % (name, msg, (mean.abs() / std).median().item(), (mean.abs() / std).max().item()))
print("MPJPE", pose.utils.compute_mpjpe(out.view(B, 17, 3), po).item())
