"""Developer: one warm NMS n = 2000 call a few times (for ncu launch lists)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200.ops import nms3d_batched  # noqa: E402

dev = torch.device("cuda:0")
d1 = torch.from_numpy(synth.c1_boxes(2000, seed=1)).to(dev)[None].contiguous()
for _ in range(4):
    out = nms3d_batched(d1, None, 0.7)
torch.cuda.synchronize()
print([o.shape for o in out if hasattr(o, "shape")])
