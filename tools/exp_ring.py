"""Streamed forward kernel after the early slot release: ring geometries (tuning key 7) and the supply / arithmetic / store
split (key 9) on C2, kernel time by CUDA events over back-to-back calls."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from roi3d_b200 import _lib
from roi3d_b200.ops import RoIAlign3D
dev = torch.device("cuda:0")
layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
f = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
r = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
def t(n=100):
    for _ in range(5): layer(f, r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): layer(f, r)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
for cfg in (int(x) for x in (sys.argv[1:] or ["0", "1", "2"])):
    _lib.set_tuning(7, cfg)
    for dbg in (0, 1, 2, 3):
        _lib.set_tuning(9, dbg)
        print("ring cfg %d debug %d (1 = no arithmetic, 2 = no store): %.1f us per step" % (cfg, dbg, t()), flush=True)
_lib.set_tuning(7, 0); _lib.set_tuning(9, 0)
