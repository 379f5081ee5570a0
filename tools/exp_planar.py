"""Developer timings of the planar forward kernel: C2 (NCDHW native / channels-last) and C3 (both layouts)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200 import SingleRoIExtractor, _lib, _util  # noqa: E402
from roi3d_b200.ops import RoIAlign3D  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
smem_opts = [int(a) for a in sys.argv[1:]] or [0]


def timeit(fn, iters=9, warm=2):
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


layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
big = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
r_big = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
big_nc = big.contiguous()
ref = layer(big, r_big)
print("C2 channels-last streamed: %.1f us" % timeit(lambda: layer(big, r_big)), flush=True)
print("C2 NCDHW convert + streamed: %.1f us" % timeit(lambda: layer(big_nc, r_big)), flush=True)
_util.FORCE_NATIVE_NCDHW[0] = True
for sm in smem_opts:
    _lib.set_tuning(10, sm)
    got = layer(big_nc, r_big)
    print("[smem %d] C2 NCDHW planar: %.1f us  maxdiff vs streamed %g" % (sm, timeit(lambda: layer(big_nc, r_big)), float((got - ref).abs().max())), flush=True)
    _lib.set_tuning(0, 60)
    got = layer(big, r_big)
    print("[smem %d] C2 channels-last planar (v60): %.1f us  maxdiff %g" % (sm, timeit(lambda: layer(big, r_big)), float((got - ref).abs().max())), flush=True)
    _lib.set_tuning(0, 0)
del big_nc, got, ref, big
dims = [(40, 128, 128), (20, 64, 64), (10, 32, 32), (5, 16, 16)]
gen = torch.Generator(device=dev); gen.manual_seed(3)
pyr_nc = [torch.randn((2, 256) + d, device=dev, generator=gen) for d in dims]
pyr_cl = [t.contiguous(memory_format=torch.channels_last_3d) for t in pyr_nc]
r3 = torch.from_numpy(synth.c3_rois(512, vols=2, seed=4)).to(dev)
ext = SingleRoIExtractor(dict(type='RoIAlign3D', out_size=14, out_size_depth=14, sample_num=2), 256, [4, 8, 16, 32], [2, 4, 8, 16])
_lib.set_tuning(0, 50)
a = ext(pyr_cl, r3)
print("C3 channels-last ring2 (v50): %.1f us" % timeit(lambda: ext(pyr_cl, r3), iters=5), flush=True)
_lib.set_tuning(0, 0)
for sm in smem_opts:
    _lib.set_tuning(10, sm)
    b = ext(pyr_cl, r3)
    print("[smem %d] C3 channels-last planar: %.1f us  maxdiff vs ring2 %g" % (sm, timeit(lambda: ext(pyr_cl, r3), iters=5), float((a - b).abs().max())), flush=True)
    c = ext(pyr_nc, r3)
    print("[smem %d] C3 NCDHW planar: %.1f us  maxdiff %g" % (sm, timeit(lambda: ext(pyr_nc, r3), iters=5), float((a - c).abs().max())), flush=True)
_lib.set_tuning(10, 0)
