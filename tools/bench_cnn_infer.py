"""Quick timing of the eval-mode CNN forward (BASELINE config 5 sweep) -- development aid.
    python tools/bench_cnn_infer.py [B ...]
"""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pose = importlib.import_module("3dhumanposeestimation_b200")
dev = torch.device("cuda", 0)
cfg = pose.ModelConfig("cnn", image_size=(256, 256), heatmap_size=256)
torch.manual_seed(0)
m = pose.CNNPoseEstimation(cfg).to(dev).eval()
for B in [int(a) for a in sys.argv[1:]] or [1, 8, 32, 128]:
    img, dep = torch.rand(B, 3, 256, 256, device=dev), torch.rand(B, 1, 256, 256, device=dev)
    kp = torch.rand(B, 17, 2, device=dev) * 0.9 + 0.05
    with torch.no_grad():
        for _ in range(3):
            m(img, dep, kp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            m(img, dep, kp)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / n * 1e3
    ms = e0.elapsed_time(e1) / n
    plan = m._plans[(B, 0)]
    print(f"B={B:5d}  {ms:8.3f} ms/fwd (wall {wall:.3f})  {B / ms * 1e3:10.1f} samples/s  "
          f"{16.559 * B / ms:8.1f} TFLOP/s  launches {plan.launches + 1}")
