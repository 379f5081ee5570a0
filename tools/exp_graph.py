"""Developer experiment: GPU-only time of the C2 forward step (plan kernel + streamed kernel) replayed as a CUDA graph."""
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
layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
f = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
r = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
for _ in range(3):
    out = layer(f, r)
torch.cuda.synchronize()
N = 20
for dbg, name in ((0, "plan + streamed kernel"), (16, "plan kernel only")):
    _lib.set_tuning(9, dbg)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            layer(f, r)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                out = layer(f, r)
    torch.cuda.synchronize()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.2f us per step (graph of %d steps)" % (name, e0.elapsed_time(e1) * 1e3 / (5 * N), N), flush=True)
_lib.set_tuning(9, 0)
