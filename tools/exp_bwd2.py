"""Developer: C2 backward timing (streamed backward kernel, zero-fill included) beside the forward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
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
x = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
r = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
out = layer(x, r)
g = torch.randn_like(out)


def bwd():
    x.grad = None
    out.backward(g, retain_graph=True)


print("C2 backward incl. zero-fill: %.1f us" % timeit(bwd), flush=True)
z = torch.empty_like(x)
print("zero-fill alone: %.1f us" % timeit(lambda: z.zero_()), flush=True)
