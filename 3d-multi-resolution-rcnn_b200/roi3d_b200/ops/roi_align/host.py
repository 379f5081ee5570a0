"""Host-buffer form of RoIAlign3D forward: numpy in, numpy out, every copy inside the call.

This is the C-ABI entry `roi3d_roi_align3d_forward_host` (include/roi3d_b200.h) -- what a non-PyTorch caller of
the reference's `roi_align_cuda.forward3d` (mmdet/ops/roi_align/src/roi_align_cuda.cpp:89-113) would bind, and
the call bench.py's `e2e` leg times.  Large single volumes are pipelined in z slabs (H2D, layout conversion,
kernel and D2H overlap); the result is identical either way.
"""
import ctypes

import numpy as np

from ... import _lib


def roi_align_3d_host(features, rois, out_size, out_size_depth, spatial_scale, spatial_scale_depth, sample_num=0,
                      layout="NCDHW", out=None):
    """features: float32 [B,C,D,H,W] (layout "NCDHW") or [B,D,H,W,C] ("NDHWC"), C-contiguous numpy array;
    rois: float32 [K,7] (batch, x1, y1, x2, y2, z1, z2).  Returns float32 [K,C,out_size_depth,out_size,out_size]."""
    if isinstance(out_size, (tuple, list)):
        ph, pw = int(out_size[0]), int(out_size[1])
    else:
        ph = pw = int(out_size)
    f = np.ascontiguousarray(features, dtype=np.float32)
    r = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 7)
    if layout == "NCDHW":
        B, C, D, H, W = f.shape
        lay = _lib.NCDHW
    elif layout == "NDHWC":
        B, D, H, W, C = f.shape
        lay = _lib.NDHWC
    else:
        raise ValueError("layout must be 'NCDHW' or 'NDHWC'")
    K = r.shape[0]
    shape = (K, C, int(out_size_depth), ph, pw)
    if out is None:
        out = np.empty(shape, np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == shape
    _lib.check(_lib.lib.roi3d_roi_align3d_forward_host(
        f.ctypes.data_as(ctypes.c_void_p), lay, B, C, D, H, W, r.ctypes.data_as(ctypes.c_void_p), K,
        int(out_size_depth), ph, pw, float(spatial_scale), float(spatial_scale_depth), int(sample_num),
        out.ctypes.data_as(ctypes.c_void_p)))
    return out
