"""Developer timings of the C3 (mask branch) backward: planar backward vs the per-warp kernel (tuning key 1 = 50)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200 import SingleRoIExtractor, _lib  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


dims = [(40, 128, 128), (20, 64, 64), (10, 32, 32), (5, 16, 16)]
gen = torch.Generator(device=dev); gen.manual_seed(3)
pyr = [torch.randn((2, 256) + d, device=dev, generator=gen).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
       for d in dims]
r3 = torch.from_numpy(synth.c3_rois(512, vols=2, seed=4)).to(dev)
ext = SingleRoIExtractor(dict(type='RoIAlign3D', out_size=14, out_size_depth=14, sample_num=2), 256, [4, 8, 16, 32], [2, 4, 8, 16])
o3 = ext(pyr, r3)
g3 = torch.randn_like(o3)


def bwd():
    for t in pyr:
        t.grad = None
    o3.backward(g3, retain_graph=True)


res = {}
for name, v in (("per-warp bwd2 (variant 50)", 50), ("planar", 0)):
    _lib.set_tuning(1, v)
    bwd()
    res[name] = [t.grad.clone() for t in pyr]
    print("C3 backward incl. zero-fill, %s: %.1f us" % (name, timeit(bwd)), flush=True)
_lib.set_tuning(1, 0)
a, b = res["per-warp bwd2 (variant 50)"], res["planar"]
for l in range(4):
    d = (a[l] - b[l]).abs().max().item()
    print("level %d: max |diff| %.3g  (max |grad| %.3g)" % (l, d, a[l].abs().max().item()))

# ---- C2 (bbox branch, 7^3): per-warp bwd2 (default) vs the planar backward (variant 60)
from roi3d_b200.ops import RoIAlign3D  # noqa: E402
del pyr, o3, g3, res, a, b
torch.cuda.empty_cache()
f = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
r2 = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
o2 = layer(f, r2)
g2 = torch.randn_like(o2)


def bwd2():
    f.grad = None
    o2.backward(g2, retain_graph=True)


grads = {}
for name, v in (("per-warp bwd2", 50), ("planar (variant 60)", 60), ("streamed", 0)):
    _lib.set_tuning(1, v)
    bwd2()
    grads[name] = f.grad.clone()
    print("C2 backward incl. zero-fill, %s: %.1f us" % (name, timeit(bwd2)), flush=True)
_lib.set_tuning(1, 0)
print("C2 max |diff| planar %.3g streamed %.3g" % ((grads["per-warp bwd2"] - grads["planar (variant 60)"]).abs().max().item(), (grads["per-warp bwd2"] - grads["streamed"]).abs().max().item()))
