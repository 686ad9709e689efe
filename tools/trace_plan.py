"""Per-launch timing of one training step: wraps the launch plan's `call` with CUDA events and prints every C-ABI call with
its integer arguments (shapes) and its device time, so that bytes / flops per launch can be set against the time."""
import argparse, importlib, os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
train = importlib.import_module("3dhumanposeestimation_b200.train")
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="cnn"); ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--top", type=int, default=70)
a = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(0)
if a.model == "vit":
    model = pose.TransformerPoseEstimation(pose.ModelConfig("transformer", image_size=(256, 256), vit_pretrained=False)).to(dev).train()
else:
    model = pose.CNNPoseEstimation(pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256)).to(dev).train()
B = a.batch
img, dep = torch.rand(B, 3, 256, 256, device=dev), torch.rand(B, 1, 256, 256, device=dev)
kp = torch.rand(B, 17, 2, device=dev) * 0.9 + 0.05
gt = torch.randn(B, 17, 3, device=dev) * 300
tr = train.Trainer(model, pose.ComprehensivePoseLoss(), lr=1e-4)
for _ in range(3):
    tr.step(img, dep, kp, gt)
torch.cuda.synchronize()
plan = model.plan(B, dev)
log = []
orig = type(plan).call
def traced(self, name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(self, name, *args)
    e1.record()
    ints = tuple(x for x in args if isinstance(x, int) and abs(x) < (1 << 31))
    log.append((name, ints, e0, e1))
type(plan).call = traced
tr.step(img, dep, kp, gt)
torch.cuda.synchronize()
type(plan).call = orig
agg = {}
for name, ints, e0, e1 in log:
    k = (name, ints)
    t = e0.elapsed_time(e1) * 1e3
    c = agg.setdefault(k, [0, 0.0])
    c[0] += 1; c[1] += t
tot = sum(v[1] for v in agg.values())
print(f"{len(log)} launches, {tot / 1e3:.2f} ms (sum of per-launch event times)")
for (name, ints), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"{t:9.1f} us {100 * t / tot:5.1f}%  x{n:<3d} {t / n:8.1f} us  {name} {ints}")
