"""Developer probe: PCIe rates and the host-buffer forward (plain vs pipelined) on the C2 workload."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
f = torch.randn(1, 256, 40, 128, 128).pin_memory()
rois = torch.from_numpy(synth.c2_rois(512, seed=2)).pin_memory()
out = torch.empty(512, 256, 7, 7, 7).pin_memory()
fd = torch.empty_like(f, device=dev)
od = torch.empty_like(out, device=dev)


def wall(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("H2D 671MB ms", wall(lambda: fd.copy_(f, non_blocking=True)))
print("D2H 180MB ms", wall(lambda: out.copy_(od, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        fd.copy_(f, non_blocking=True)
    with torch.cuda.stream(s2):
        out.copy_(od, non_blocking=True)


print("H2D + D2H concurrently ms", wall(both))


def host_call():
    _lib.check(_lib.lib.roi3d_roi_align3d_forward_host(
        ctypes.c_void_p(f.data_ptr()), _lib.NCDHW, 1, 256, 40, 128, 128, ctypes.c_void_p(rois.data_ptr()), 512,
        7, 7, 7, 0.25, 0.5, 2, ctypes.c_void_p(out.data_ptr())))


for kb in (-1, 0):
    _lib.set_tuning(4, kb)
    print("host forward, pipeline_kb", kb, "ms", wall(host_call))
_lib.set_tuning(4, 0)

fcl = f.permute(0, 2, 3, 4, 1).contiguous().pin_memory()


def host_call_cl():
    _lib.check(_lib.lib.roi3d_roi_align3d_forward_host(
        ctypes.c_void_p(fcl.data_ptr()), _lib.NDHWC, 1, 256, 40, 128, 128, ctypes.c_void_p(rois.data_ptr()), 512,
        7, 7, 7, 0.25, 0.5, 2, ctypes.c_void_p(out.data_ptr())))


for kb in (-1, 0):
    _lib.set_tuning(4, kb)
    print("host forward NDHWC, pipeline_kb", kb, "ms", wall(host_call_cl))
_lib.set_tuning(4, 0)
# group sizes of the pipelined schedule
r = rois.numpy()
last = np.minimum(39, np.floor((np.maximum(r[:, 5], r[:, 6]) + 1) * 0.5) + 2)
print("RoIs per slab group", np.bincount((last // 4).astype(int), minlength=10))
