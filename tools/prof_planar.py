"""Run the planar forward kernel a few times (for ncu).  usage: prof_planar.py {c2|c3}"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200 import SingleRoIExtractor, _util  # noqa: E402
_util.FORCE_NATIVE_NCDHW[0] = True
from roi3d_b200.ops import RoIAlign3D  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "c2"
if what == "c2":
    f = torch.randn(1, 256, 40, 128, 128, device=dev)
    rois = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
    layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
    for _ in range(3):
        layer(f, rois)
else:
    dims = [(40, 128, 128), (20, 64, 64), (10, 32, 32), (5, 16, 16)]
    pyr = [torch.randn((2, 256) + d, device=dev) for d in dims]
    rois = torch.from_numpy(synth.c3_rois(512, vols=2, seed=4)).to(dev)
    ex = SingleRoIExtractor(dict(type='RoIAlign3D', out_size=14, out_size_depth=14, sample_num=2), 256,
                            [4, 8, 16, 32], [2, 4, 8, 16])
    for _ in range(3):
        ex(pyr, rois)
torch.cuda.synchronize()
print("done")
