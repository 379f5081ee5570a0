"""Developer experiment: DRAM traffic of the NCDHW twin under different item schedules (run under ncu)."""
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
nat = torch.randn(1, 256, 40, 128, 128, device=dev)
r = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
for dbg in [int(a) for a in sys.argv[1:]] or [0]:
    _lib.set_tuning(9, dbg)
    for _ in range(3):
        out = layer(nat, r)
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); layer(nat, r); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("debug %d: %.1f us" % (dbg, sorted(ts)[4]), flush=True)
_lib.set_tuning(9, 0)
