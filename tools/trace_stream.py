"""Developer trace of CTA 0 of the streamed forward kernel: per-tile clocks of the producer and of owner warp 0."""
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
layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
f = torch.randn(1, 256, 40, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
r = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
trace = torch.zeros(4 * 8192, dtype=torch.int64, device=dev)
p = trace.data_ptr()
for dbg in (int(x) for x in (sys.argv[1:] or ["0", "3"])):
    _lib.set_tuning(9, dbg)
    for _ in range(2):
        layer(f, r)
    trace.zero_()
    lo, hi = p & 0xffffffff, p >> 32
    _lib.set_tuning(10, lo - (1 << 32) if lo >= (1 << 31) else lo)
    _lib.set_tuning(11, hi)
    layer(f, r)
    torch.cuda.synchronize()
    _lib.set_tuning(10, 0)
    _lib.set_tuning(11, 0)
    t = trace.cpu().numpy().reshape(4, 8192)
    n = int((t[0] != 0).sum())
    pw, pi_raw, of, od = t[0, :n], t[1, :n], t[2, :n], t[3, :n]
    pi = pi_raw & ((1 << 48) - 1)
    rows = (pi_raw >> 48) & 0xfff
    first = (pi_raw >> 60) & 1
    t0 = pw[0]
    print("debug=%d tiles=%d items=%d total cycles=%d  cycles/tile=%.0f" % (dbg, n, int(first.sum()), od[-1] - t0, (od[-1] - t0) / n))
    print(" producer: issue cost (after empty wait -> ops issued) mean %.0f" % np.mean(pi - pw))
    print(" producer: gap between consecutive tile issues mean %.0f median %.0f" % (np.mean(np.diff(pw)), np.median(np.diff(pw))))
    print(" owner: full-wait done minus producer issue (load latency) mean %.0f median %.0f min %.0f" % (np.mean(of - pi), np.median(of - pi), np.min(of - pi)))
    print(" owner: processing (after wait -> done) mean %.0f median %.0f" % (np.mean(od - of), np.median(od - of)))
    print(" owner: gap done(t) -> full-wait done(t+1) (stall for next tile) mean %.0f median %.0f" % (np.mean(of[1:] - od[:-1]), np.median(of[1:] - od[:-1])))
    print(" producer lead: how many tiles issued before owner finished tile t (mean): %.2f" % np.mean([np.searchsorted(pi, od[i]) - i for i in range(n)]))
    k = min(n, 40)
    print(" first %d tiles: rows, first, prod_wait_done, prod_issued, owner_got, owner_done (cycles from start)" % k)
    for i in range(k):
        print("  %3d rows %2d first %d  %7d %7d %7d %7d" % (i, rows[i], first[i], pw[i] - t0, pi[i] - t0, of[i] - t0, od[i] - t0))
_lib.set_tuning(9, 0)
