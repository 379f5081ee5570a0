"""Developer probe: where the C4 proposal path (8 volumes) spends its wall time."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
from roi3d_b200 import RPNProposal3D  # noqa: E402

dev = torch.device("cuda:0")
Bv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dims4 = [(80, 128, 128), (40, 64, 64), (20, 32, 32), (10, 16, 16), (5, 8, 8)]
gen = torch.Generator(device=dev)
gen.manual_seed(6)
cls = [2 * torch.randn((Bv, 1) + d, device=dev, generator=gen) for d in dims4]
reg = [0.1 * torch.randn((Bv, 6) + d, device=dev, generator=gen) for d in dims4]
head = RPNProposal3D(anchor_scales=[2], anchor_depth_scales=[2], anchor_ratios=[1.0],
                     anchor_strides=[4, 8, 16, 32, 64], anchor_strides_depth=[2, 4, 8, 16, 32])
cfg = dict(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7, min_bbox_size=0, nms_across_levels=False)
metas = [dict(img_shape=(512, 512, 3, 160), scale_factor=1.0)] * Bv
for _ in range(3):
    head.get_proposals(cls, reg, metas, cfg)
torch.cuda.synchronize()
for rep in range(3):
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        out = head.get_proposals(cls, reg, metas, cfg)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    print("wall per call us", sum(ts) / len(ts), "median", sorted(ts)[5], "max", max(ts), "props", [int(o.shape[0]) for o in out][:3])
head.cuda_graph = True
for _ in range(3):
    head.get_proposals(cls, reg, metas, cfg)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    out = head.get_proposals(cls, reg, metas, cfg)
torch.cuda.synchronize()
print("wall per call us (CUDA graph replay)", (time.perf_counter() - t0) / 10 * 1e6)
head.cuda_graph = False
if len(sys.argv) > 2:
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        head.get_proposals(cls, reg, metas, cfg)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
