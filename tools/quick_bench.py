"""Developer timing script (not the contract bench): times the hot-path kernels with CUDA events."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-multi-resolution-rcnn_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import synth  # noqa: E402
import roi3d_b200  # noqa: E402
from roi3d_b200 import _lib  # noqa: E402
from roi3d_b200.ops import RoIAlign3D, nms  # noqa: E402

dev = torch.device("cuda:0")
flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3, flush=True):
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
    return ts[len(ts) // 2], ts[0]


class _Res(dict):
    def __setitem__(self, k, v):
        print(k, v, flush=True)
        dict.__setitem__(self, k, v)


res = _Res()
which = sys.argv[1:] or ["c2", "nms", "c3", "ref"]
BWD_VARIANTS = [int(x) for x in os.environ.get("QB_BWD", "0,1,2").split(",")]
FWD_VARIANTS = [int(x) for x in os.environ.get("QB_FWD", "0,30").split(",")]

if "c2" in which:
    f = torch.randn(1, 256, 40, 128, 128, device=dev)
    fcl = f.contiguous(memory_format=torch.channels_last_3d)
    rois = torch.from_numpy(synth.c2_rois(512, seed=2)).to(dev)
    layer = RoIAlign3D(7, 7, 0.25, 0.5, 2)
    for v in FWD_VARIANTS:
        _lib.set_tuning(0, v % 1000)
        _lib.set_tuning(2, v // 1000)
        med, mn = timeit(lambda: layer(fcl, rois))
        res["c2_fwd_cl_v%d_us" % v] = (med, mn)
    _lib.set_tuning(0, 0)
    _lib.set_tuning(2, 0)
    med, mn = timeit(lambda: layer(fcl, rois), flush=False)
    res["c2_fwd_cl_noflush_us"] = (med, mn)
    roi3d_b200._util._CACHE_SIZE = 0
    med, mn = timeit(lambda: layer(f, rois), iters=5)
    res["c2_fwd_ncdhw_with_transpose_us"] = (med, mn)
    # sorted rois (L2 locality experiment)
    order = torch.argsort(rois[:, 5] * 1000 + rois[:, 2])
    rs = rois[order].contiguous()
    med, mn = timeit(lambda: layer(fcl, rs))
    res["c2_fwd_cl_sortedrois_us"] = (med, mn)
    # backward
    fcl.requires_grad_(True)
    out = layer(fcl, rois)
    g = torch.randn_like(out)
    for v in BWD_VARIANTS:
        _lib.set_tuning(1, v)
        def bwd():
            fcl.grad = None
            out.backward(g, retain_graph=True)
        med, mn = timeit(bwd, iters=10)
        res["c2_bwd_cl_v%d_us(incl zero-fill)" % v] = (med, mn)
    _lib.set_tuning(1, 0)
    try:
        import ref_roi_align_cuda as ra
        ref_out = torch.zeros_like(out)
        fd = f.detach()
        med, mn = timeit(lambda: ra.forward3d(fd, rois, 7, 7, 7, 0.25, 0.5, 2, ref_out), iters=5)
        res["c2_REF_fwd_us"] = (med, mn)
        ref_g = torch.zeros_like(fd)
        med, mn = timeit(lambda: ra.backward3d(g, rois, 7, 7, 7, 0.25, 0.5, 2, ref_g), iters=3, warm=1)
        res["c2_REF_bwd_us(no zero-fill)"] = (med, mn)
        mine = layer(fcl.detach(), rois)
        res["c2_max_abs_diff_vs_ref"] = float((mine - ref_out).abs().max())
    except Exception as e:
        res["ref_roi_align_error"] = repr(e)
    del f, fcl, out, g
    torch.cuda.empty_cache()

if "nms" in which:
    dets = torch.from_numpy(synth.c1_boxes(2000, seed=0)).to(dev)
    med, mn = timeit(lambda: nms(dets, 0.7), iters=50, flush=False)
    res["nms2000_wrapper_us(incl count sync)"] = (med, mn)
    from roi3d_b200.ops import nms3d_batched
    d1 = dets.unsqueeze(0).contiguous()
    med, mn = timeit(lambda: nms3d_batched(d1, None, 0.7), iters=50, flush=False)
    res["nms2000_device_only_us"] = (med, mn)
    d40 = dets.unsqueeze(0).repeat(40, 1, 1).contiguous()
    med, mn = timeit(lambda: nms3d_batched(d40, None, 0.7), iters=20, flush=False)
    res["nms2000_x40_batched_us"] = (med, mn)
    dn = synth.c1_boxes(2000, seed=0)
    t0 = time.perf_counter()
    for _ in range(20):
        nms(dn, 0.7, device_id=0)
    res["nms2000_host_numpy_us"] = (time.perf_counter() - t0) / 20 * 1e6
    try:
        import ref_nms_cuda
        med, mn = timeit(lambda: ref_nms_cuda.nms_3d(dets, 0.7), iters=20, flush=False)
        res["nms2000_REF_us"] = (med, mn)
    except Exception as e:
        res["ref_nms_error"] = repr(e)

if "c3" in which:
    from roi3d_b200 import SingleRoIExtractor
    dims = [(40, 128, 128), (20, 64, 64), (10, 32, 32), (5, 16, 16)]
    feats = [torch.randn((2, 256) + d, device=dev).contiguous(memory_format=torch.channels_last_3d) for d in dims]
    rois = torch.from_numpy(synth.c3_rois(512, vols=2, seed=4)).to(dev)
    ex = SingleRoIExtractor(dict(type='RoIAlign3D', out_size=14, out_size_depth=14, sample_num=2), 256,
                            [4, 8, 16, 32], [2, 4, 8, 16])
    lv = ex.map_roi_levels(rois, 4)
    res["c3_level_hist"] = np.bincount(lv.cpu().numpy(), minlength=4).tolist()
    for v in FWD_VARIANTS:
        _lib.set_tuning(0, v % 1000)
        _lib.set_tuning(2, v // 1000)
        med, mn = timeit(lambda: ex(feats, rois), iters=5)
        res["c3_fwd_v%d_us" % v] = (med, mn)
    _lib.set_tuning(0, 0)
    _lib.set_tuning(2, 0)
    for f in feats:
        f.requires_grad_(True)
    out = ex(feats, rois)
    g = torch.randn_like(out)
    for v in BWD_VARIANTS:
        _lib.set_tuning(1, v)
        def bwd3():
            for f in feats:
                f.grad = None
            out.backward(g, retain_graph=True)
        med, mn = timeit(bwd3, iters=5, warm=1)
        res["c3_bwd_v%d_us(incl zero-fill)" % v] = (med, mn)
    _lib.set_tuning(1, 0)

if "c4" in which:
    from roi3d_b200 import RPNProposal3D
    from roi3d_b200.models.anchor_heads import topk_segmented
    B = 8
    dims = [(80, 128, 128), (40, 64, 64), (20, 32, 32), (10, 16, 16), (5, 8, 8)]
    g = torch.Generator(device=dev); g.manual_seed(6)
    cls = [2 * torch.randn((B, 1) + d, device=dev, generator=g) for d in dims]
    reg = [0.1 * torch.randn((B, 6) + d, device=dev, generator=g) for d in dims]
    head = RPNProposal3D(anchor_scales=[2], anchor_depth_scales=[2], anchor_ratios=[1.0],
                         anchor_strides=[4, 8, 16, 32, 64], anchor_strides_depth=[2, 4, 8, 16, 32])
    cfg = dict(nms_pre=2000, nms_post=1000, max_num=1000, nms_thr=0.7, min_bbox_size=0, nms_across_levels=False)
    metas = [dict(img_shape=(512, 512, 3, 160), scale_factor=1.0)] * B
    out = head.get_proposals(cls, reg, metas, cfg)
    res["c4_num_proposals"] = [int(o.shape[0]) for o in out]
    t0 = time.perf_counter(); torch.cuda.synchronize()
    for _ in range(5):
        head.get_proposals(cls, reg, metas, cfg)
    torch.cuda.synchronize()
    res["c4_get_bboxes_8vol_wall_us"] = (time.perf_counter() - t0) / 5 * 1e6
    segs = [cls[l][b] for b in range(B) for l in range(5)]
    med, mn = timeit(lambda: topk_segmented(segs, 2000, apply_sigmoid=True, permute_adhw=True), iters=10, flush=False)
    res["c4_topk_40seg_us"] = (med, mn)
    # the reference's per-level composition with torch ops (sigmoid + topk) for one volume, P2 level only
    def ref_topk():
        for b in range(B):
            for l in range(5):
                s = cls[l][b].permute(2, 3, 1, 0).reshape(-1).sigmoid()
                if s.numel() > 2000:
                    s.topk(2000)
    med, mn = timeit(ref_topk, iters=5, flush=False)
    res["c4_torch_sigmoid_topk_40seg_us"] = (med, mn)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w") as fh:
    json.dump(res, fh, indent=1)
