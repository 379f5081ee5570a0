"""A/B of the segmented top-k on BASELINE C4's score maps: sieve path (tuning key 11 = 0) vs digit passes only (-1)."""
import os, sys, time
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
res = {}
for mode in (0, -1, 2, 0, -1):
    roi3d_b200._lib.set_tuning(11, mode)
    for _ in range(5):
        out = topk_segmented(segs, 2000, apply_sigmoid=True, permute_adhw=True, small_in_index_order=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = topk_segmented(segs, 2000, apply_sigmoid=True, permute_adhw=True, small_in_index_order=True)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("mode", mode, "graph replay us/call", e0.elapsed_time(e1) * 1000 / 50, flush=True)
    res.setdefault(mode, out)
print("identical 0 vs -1:", all(torch.equal(a, b) for a, b in zip(res[0], res[-1])),
      "0 vs 2:", all(torch.equal(a, b) for a, b in zip(res[0], res[2])))
roi3d_b200._lib.set_tuning(11, 0)
