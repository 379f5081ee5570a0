"""Developer experiments on the streamed forward kernel: what bounds it (supply, arithmetic, store)."""
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
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=15, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
big = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
r_big = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
# same RoI sizes on a map that fits in L2 (256 x 10 x 64 x 64 = 42 MB): RoIs folded into the smaller volume
small = torch.randn(1, 256, 10, 64, 64, device=dev).contiguous(memory_format=torch.channels_last_3d)
rs = synth.c2_rois(512, seed=2).copy()
w, h, d = rs[:, 3] - rs[:, 1], rs[:, 4] - rs[:, 2], np.minimum(rs[:, 6] - rs[:, 5], 14)
rs[:, 1] = rs[:, 1] % (256 - 66); rs[:, 3] = rs[:, 1] + w
rs[:, 2] = rs[:, 2] % (256 - 66); rs[:, 4] = rs[:, 2] + h
rs[:, 5] = rs[:, 5] % 4; rs[:, 6] = rs[:, 5] + d
r_small = torch.from_numpy(rs).to(dev)
for name, f, r, fl in (("HBM-sized map", big, r_big, True), ("L2-resident map", small, r_small, False)):
    for dbg in (0, 1, 2, 3):
        _lib.set_tuning(9, dbg)
        t = timeit(lambda: layer(f, r), flush=fl)
        print("%s debug=%d (1: no arithmetic, 2: no store, 4: roi-major item order): %.1f us" % (name, dbg, t), flush=True)
    _lib.set_tuning(9, 0)
    _lib.set_tuning(0, 50)
    print("%s ring2: %.1f us" % (name, timeit(lambda: layer(f, r), flush=fl)), flush=True)
    _lib.set_tuning(0, 0)

# ---- planar kernel: NCDHW native and channels-last (variant 60), C2 and C3
print("--- planar kernel", flush=True)
big_nc = big.contiguous()   # NCDHW
ref = layer(big, r_big)
got = layer(big_nc, r_big)
print("C2 NCDHW native maxdiff vs streamed %g" % float((got - ref).abs().max()), flush=True)
print("C2 NCDHW native (planar): %.1f us" % timeit(lambda: layer(big_nc, r_big)), flush=True)
_lib.set_tuning(0, 60)
print("C2 channels-last planar (v60): %.1f us" % timeit(lambda: layer(big, r_big)), flush=True)
_lib.set_tuning(0, 0)
del big_nc, got, ref
from roi3d_b200 import SingleRoIExtractor
dims = [(40, 128, 128), (20, 64, 64), (10, 32, 32), (5, 16, 16)]
gen = torch.Generator(device=dev); gen.manual_seed(3)
pyr_nc = [torch.randn((2, 256) + d, device=dev, generator=gen) for d in dims]
pyr_cl = [t.contiguous(memory_format=torch.channels_last_3d) for t in pyr_nc]
r3 = torch.from_numpy(synth.c3_rois(512, vols=2, seed=4)).to(dev)
ext = SingleRoIExtractor(dict(type='RoIAlign3D', out_size=14, out_size_depth=14, sample_num=2), 256, [4, 8, 16, 32], [2, 4, 8, 16])
a = ext(pyr_cl, r3)
print("C3 channels-last (auto = planar): %.1f us" % timeit(lambda: ext(pyr_cl, r3), iters=5), flush=True)
_lib.set_tuning(0, 50)
b = ext(pyr_cl, r3)
print("C3 channels-last ring2 (v50): %.1f us   maxdiff %g" % (timeit(lambda: ext(pyr_cl, r3), iters=5), float((a - b).abs().max())), flush=True)
_lib.set_tuning(0, 0)
c = ext(pyr_nc, r3)
print("C3 NCDHW native: %.1f us   maxdiff %g" % (timeit(lambda: ext(pyr_nc, r3), iters=5), float((a - c).abs().max())), flush=True)
