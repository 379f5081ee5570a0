"""Developer script: localise differences between the streamed forward kernel and the ring kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200 import _lib  # noqa: E402
from roi3d_b200.ops import RoIAlign3D  # noqa: E402

dev = torch.device("cuda:0")


def run(f, rois, pdp=7, variant=0):
    _lib.set_tuning(0, variant)
    try:
        out = RoIAlign3D(7, pdp, 0.25, 0.5, 2)(f, rois)
        torch.cuda.synchronize()
    finally:
        _lib.set_tuning(0, 0)
    return out


def coord_feats(C, D, H, W, which):
    z, y, x = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing="ij")
    base = {"x": x, "y": y, "z": z, "one": torch.ones_like(x)}[which].float()
    f = base[None, None].repeat(1, C, 1, 1, 1)
    if which == "c":
        f = torch.arange(C).float()[None, :, None, None, None].repeat(1, 1, D, H, W)
    return f.to(dev).contiguous(memory_format=torch.channels_last_3d)


C, D, H, W = 64, 10, 32, 32
rois = torch.tensor([[0, 20.0, 24.0, 60.0, 70.0, 4.0, 12.0],
                     [0, 10.0, 10.0, 10.0, 10.0, 4.0, 4.0],
                     [0, 3.25, 7.75, 41.5, 29.125, 2.5, 17.75]], device=dev)
for which in ["one", "x", "y", "z"]:
    f = coord_feats(C, D, H, W, which)
    a = run(f, rois, variant=0)
    b = run(f, rois, variant=50)
    diff = (a - b).abs()
    print("feat=%s maxdiff=%g" % (which, float(diff.max())))
    if float(diff.max()) > 1e-4:
        for k in range(rois.shape[0]):
            dk = diff[k]
            print("  roi %d: maxdiff %g; per-channel max (first 8): %s" % (
                k, float(dk.max()), np.round(dk.amax(dim=(1, 2, 3))[:8].cpu().numpy(), 4)))
            print("   stream c0 pd0:\n", np.round(a[k, 0, 0].cpu().numpy(), 3))
            print("   ring   c0 pd0:\n", np.round(b[k, 0, 0].cpu().numpy(), 3))
            print("   stream c0 [:,0,0] over pd:", np.round(a[k, 0, :, 0, 0].cpu().numpy(), 3),
                  " ring:", np.round(b[k, 0, :, 0, 0].cpu().numpy(), 3))
            print("   stream c1 pd0 row0:", np.round(a[k, 1, 0, 0].cpu().numpy(), 3), " c63:", np.round(a[k, 63, 0, 0].cpu().numpy(), 3))

# multi-item per CTA: many RoIs
print("--- many rois")
f = torch.randn(1, 64, 10, 32, 32, device=dev).contiguous(memory_format=torch.channels_last_3d)
for K in (100, 400, 1200):
    r = torch.from_numpy(synth.c2_rois(K, seed=3, img=(128, 128, 20))).to(dev)
    a = run(f, r, variant=0)
    b = run(f, r, variant=50)
    d = (a - b).abs().amax(dim=(1, 2, 3, 4))
    print("K=%d maxdiff=%g bad rois=%d first bad=%s" % (K, float(d.max()), int((d > 1e-4).sum()), torch.nonzero(d > 1e-4)[:10].flatten().tolist()))

print("--- c2 scale", flush=True)
f = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
r = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
for K in (8, 64, 512):
    a = run(f, r[:K].contiguous(), variant=0)
    b = run(f, r[:K].contiguous(), variant=50)
    d = (a - b).abs().amax(dim=(1, 2, 3, 4))
    print("K=%d maxdiff=%g bad rois=%d" % (K, float(d.max()), int((d > 1e-4).sum())), flush=True)
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3):
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
    return ts[len(ts) // 2], ts[0]


layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
for cfg, pf in ((0, 0), (4, 0), (6, 0), (5, 0)):
    _lib.set_tuning(7, cfg)
    _lib.set_tuning(8, pf)
    a = layer(f, r)
    print("cfg %d prefetch %d maxdiff vs ring %g" % (cfg, pf, float((a - b).abs().max())), flush=True)
    print("stream cfg %d prefetch %d: median %.1f us min %.1f us (events, L2 flushed)" % ((cfg, pf) + timeit(lambda: layer(f, r))), flush=True)
_lib.set_tuning(7, 0)
_lib.set_tuning(8, 1)
_lib.set_tuning(0, 50)
print("ring2: median %.1f us min %.1f us" % timeit(lambda: layer(f, r)), flush=True)
_lib.set_tuning(0, 0)
