"""Times the training step of a model on one GPU (CUDA events), optionally printing a per-kernel breakdown."""
import argparse, importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pose = importlib.import_module("3dhumanposeestimation_b200")
train = importlib.import_module("3dhumanposeestimation_b200.train")
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="vit"); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=5); ap.add_argument("--profile", action="store_true")
ap.add_argument("--warmup", type=int, default=3); ap.add_argument("--dropout0", action="store_true")
ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
a = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(0)
if a.model == "vit":
    kw = dict(transformer_dropout_rate=0.0, transformer_attention_dropout_rate=0.0, regression_dropout=0.0) if a.dropout0 else {}
    cfg = pose.ModelConfig("transformer", image_size=(256, 256), vit_pretrained=False, **kw)
    model = pose.TransformerPoseEstimation(cfg).to(dev).train()
else:
    cfg = pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256)
    model = pose.CNNPoseEstimation(cfg).to(dev).train()
B = a.batch
img, dep = torch.rand(B, 3, 256, 256, device=dev), torch.rand(B, 1, 256, 256, device=dev)
kp = torch.rand(B, 17, 2, device=dev) * 0.9 + 0.05
gt = torch.randn(B, 17, 3, device=dev) * 300
tr = train.Trainer(model, pose.ComprehensivePoseLoss(), lr=1e-4)
for _ in range(a.warmup):
    o5 = tr.step(img, dep, kp, gt)
torch.cuda.synchronize()
print("loss", o5.tolist())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if a.cuda_profiler:
    torch.cuda.cudart().cudaProfilerStart()
t0 = time.perf_counter(); e0.record()
for _ in range(a.steps):
    o5 = tr.step(img, dep, kp, gt)
e1.record(); torch.cuda.synchronize()
if a.cuda_profiler:
    torch.cuda.cudart().cudaProfilerStop()
ms = e0.elapsed_time(e1) / a.steps
print(f"{a.model} B={B}: {ms:.2f} ms/step  {B / ms * 1e3:.1f} samples/s  wall {(time.perf_counter() - t0) / a.steps * 1e3:.2f} ms  loss {o5[4].item():.3f}")
plan = model.plan(B, dev)
print("launches/step (fwd+bwd)", plan.launches)
if a.profile:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.step(img, dep, kp, gt); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
