"""Small driver for ncu: runs each hot-path kernel a few times at the bench sizes (B=256, 256x256).
    python tools/prof_kernels.py [augment|heatmap|loss|head|all] [reps]
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pose = importlib.import_module("3dhumanposeestimation_b200")
dev = torch.device("cuda", 0)
inp = bench.make_inputs(0)
d = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
aug = pose.PoseAugmentor()
params = bench.draw_params(aug, bench.B)
hm = pose.GaussianHeatmapGenerator(bench.J, bench.HS, bench.SIGMA).to(dev)
head = pose.PoseRegressionHead(bench.HEAD_IN, bench.J, hidden_dims=list(bench.HEAD_HIDDEN), activation="silu").to(dev).eval()
for _ in range(reps):
    if which in ("augment", "all"):
        a = aug.augment_batch(d["image"], d["depth"], d["kp"], d["joints"], d["cam"], params=params, pad_to=(308, 308))
    if which in ("heatmap", "all"):
        hm(d["kp"])
    if which in ("head", "all"):
        with torch.no_grad():
            pred = head(d["feat"])
    if which in ("loss", "all"):
        pose.loss.pose_loss_fwd_bwd(d["joints"], d["joints"] + 1.0, (1.0, 1.0, 100.0, 1.0))
torch.cuda.synchronize()
print("ok")
