import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-multi-resolution-rcnn_b200"))
import torch
import roi3d_b200
from roi3d_b200.models.anchor_heads import topk_segmented
dev = torch.device("cuda:0")
dims = [(80, 128, 128), (40, 64, 64), (20, 32, 32), (10, 16, 16), (5, 8, 8)]
gen = torch.Generator(device=dev)
gen.manual_seed(6)
cls = [2 * torch.randn((8, 1) + d, device=dev, generator=gen) for d in dims]
segs = [cls[l][b] for b in range(8) for l in range(5)]
for mode in (0, 0, 0, -1):
    roi3d_b200._lib.set_tuning(11, mode)
    out = topk_segmented(segs, 2000, apply_sigmoid=True, permute_adhw=True, small_in_index_order=True)
torch.cuda.synchronize()
