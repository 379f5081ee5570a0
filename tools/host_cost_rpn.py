"""Host cost of the eager C4 proposal path: enqueue-only wall time per call and the Python profile of one call."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "3d-multi-resolution-rcnn_b200"))
import torch
from roi3d_b200 import RPNProposal3D
dev = torch.device("cuda:0")
Bv = 8
dims4 = [(80, 128, 128), (40, 64, 64), (20, 32, 32), (10, 16, 16), (5, 8, 8)]
gen = torch.Generator(device=dev); gen.manual_seed(6)
cls = [2 * torch.randn((Bv, 1) + d, device=dev, generator=gen) for d in dims4]
reg = [0.1 * torch.randn((Bv, 6) + d, device=dev, generator=gen) for d in dims4]
head = RPNProposal3D(anchor_scales=[2], anchor_depth_scales=[2], anchor_ratios=[1.0],
                     anchor_strides=[4, 8, 16, 32, 64], anchor_strides_depth=[2, 4, 8, 16, 32])
cfg = dict(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7, min_bbox_size=0, nms_across_levels=False)
metas = [dict(img_shape=(512, 512, 3, 160), scale_factor=1.0)] * Bv
for _ in range(5):
    head.get_proposals(cls, reg, metas, cfg)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(100):
        head._enqueue(cls, reg, metas, cfg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("enqueue-only per call us %.1f   incl. drain %.1f" % ((t1 - t0) * 1e4, (t2 - t0) * 1e4))
for rep in range(2):
    t0 = time.perf_counter()
    for _ in range(100):
        head.get_proposals(cls, reg, metas, cfg)
    print("get_proposals (host read per call) us %.1f" % ((time.perf_counter() - t0) * 1e4))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    head._enqueue(cls, reg, metas, cfg)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
