import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-multi-resolution-rcnn_b200"))
import numpy as np, torch
import roi3d_b200
from roi3d_b200.models.anchor_heads import topk_segmented
dev = torch.device("cuda:0")
rng = np.random.default_rng(77)
w = rng.standard_normal(50000).astype(np.float32)
w[::997] = np.nan
t = torch.from_numpy(w).to(dev)
print("nan on device", int(torch.isnan(t).sum()))
for mode in (0, -1, 2):
    roi3d_b200._lib.set_tuning(11, mode)
    idx, val = topk_segmented([t], 2000)
    print(mode, idx[0, :5].tolist(), val[0, :5].tolist(), int(torch.isnan(val).sum()))
tv, ti = torch.topk(t, 2000)
print("torch", ti[:5].tolist(), tv[:5].tolist())
