"""Developer: host (CPU) cost per call of the small ops' Python / ctypes binding.
CPU time is taken over batches of calls that fit the launch queue, with the GPU drained between batches."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from roi3d_b200 import _lib  # noqa: E402
from roi3d_b200._util import stream_ptr, workspace  # noqa: E402
from roi3d_b200.ops import nms3d_batched  # noqa: E402

dev = torch.device("cuda:0")


def cpu_us(fn, calls=100, rounds=7):
    ts = []
    for _ in range(rounds):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(calls):
            fn()
        ts.append((time.perf_counter() - t0) / calls * 1e6)
        torch.cuda.synchronize()
    ts.sort()
    return ts[len(ts) // 2]


d1 = torch.from_numpy(synth.c1_boxes(2000, seed=0)).to(dev)[None].contiguous()
print("nms3d_batched (python wrapper), CPU us per call: %.1f" % cpu_us(lambda: nms3d_batched(d1, None, 0.7)))
keep = torch.empty((1, 2000), dtype=torch.int64, device=dev)
keep_s = torch.empty_like(keep)
num = torch.empty((1,), dtype=torch.int32, device=dev)
nbytes = _lib.lib.roi3d_nms3d_workspace_bytes(1, 2000)
buf, ws = workspace(dev, nbytes)
sp = stream_ptr()
args = (d1.data_ptr(), None, None, 1, 2000, 0.7, keep.data_ptr(), keep_s.data_ptr(), num.data_ptr(), ws, nbytes, sp)
print("roi3d_nms3d_batched_presorted through ctypes alone (3 launches): %.1f" % cpu_us(lambda: _lib.lib.roi3d_nms3d_batched_presorted(*args)))
print("stream_ptr(): %.1f" % cpu_us(stream_ptr, calls=1000))


def ctx():
    with torch.cuda.device(dev):
        pass


print("with torch.cuda.device(dev): %.1f" % cpu_us(ctx, calls=1000))
print("4 x torch.empty: %.1f" % cpu_us(lambda: (torch.empty((1, 2000), dtype=torch.int64, device=dev), torch.empty((1, 2000), dtype=torch.int64, device=dev),
                                                 torch.empty((1,), dtype=torch.int32, device=dev), torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)), calls=1000))
# GPU-side: back-to-back calls, GPU never idle
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    nms3d_batched(d1, None, 0.7)
e1.record()
torch.cuda.synchronize()
print("200 back-to-back wrapper calls: %.1f us per call on the device" % (e0.elapsed_time(e1) * 1e3 / 200))
