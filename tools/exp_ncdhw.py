"""Developer experiment: the NCDHW twin of the streamed forward kernel on C2 (supply / arithmetic / store split)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200 import _lib  # noqa: E402
from roi3d_b200.ops import RoIAlign3D  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=15, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
nat = torch.randn(1, 256, 40, 128, 128, device=dev)
cl = nat.contiguous(memory_format=torch.channels_last_3d)
r = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
a, b = layer(cl, r), layer(nat, r)
print("max abs diff to channels-last:", float((a - b).abs().max()), flush=True)
print("channels-last streamed: %.1f us" % timeit(lambda: layer(cl, r)), flush=True)
for dbg in (0, 16, 3, 19):
    _lib.set_tuning(9, dbg)
    print("NCDHW streamed debug=%d (1: no arithmetic, 2: no store, 16: cost order instead of Morton): %.1f us" % (dbg, timeit(lambda: layer(nat, r))), flush=True)
_lib.set_tuning(9, 0)
_lib.set_tuning(0, 60)
print("NCDHW planar: %.1f us" % timeit(lambda: layer(nat, r)), flush=True)
_lib.set_tuning(0, 0)
