"""3D box helpers of the proposal path (reference: mmdet/core/bbox/transforms.py)."""
import ctypes

import numpy as np
import torch

from ... import _lib
from ..._util import check_cuda_f32, stream_ptr


def bbox2roi3D(bbox_list):
    """[n_i, >=6] boxes per image -> [K,7] rois (batch_ind, x1, y1, x2, y2, z1, z2).
    Reference: bbox2roi3D, transforms.py:220-239."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :6]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 7))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


def delta2bbox3D(rois, deltas, means=(0, 0, 0, 0, 0, 0), stds=(1, 1, 1, 1, 1, 1), max_shape=None,
                 wh_ratio_clip=16 / 1000, d_ratio_clip=64 / 1000):
    """Torch-op form of delta2bbox3D (transforms.py:105-160) for callers outside the fused proposal path
    (e.g. the bbox head).  Note the reference clamps all of dw, dh, dz, dd with |log(wh_ratio_clip)| (:122-128);
    d_ratio_clip is accepted and unused, as there.  max_shape = img_shape (H, W, 3, D)."""
    means = deltas.new_tensor(means).repeat(1, deltas.size(1) // 6)
    stds = deltas.new_tensor(stds).repeat(1, deltas.size(1) // 6)
    d = deltas * stds + means
    dx, dy, dw, dh, dz, dd = (d[:, i::6] for i in range(6))
    max_ratio = float(np.abs(np.log(wh_ratio_clip)))
    dw, dh = dw.clamp(-max_ratio, max_ratio), dh.clamp(-max_ratio, max_ratio)
    dz, dd = dz.clamp(-max_ratio, max_ratio), dd.clamp(-max_ratio, max_ratio)
    px = ((rois[:, 0] + rois[:, 2]) * 0.5).unsqueeze(1).expand_as(dx)
    py = ((rois[:, 1] + rois[:, 3]) * 0.5).unsqueeze(1).expand_as(dy)
    pz = ((rois[:, 4] + rois[:, 5]) * 0.5).unsqueeze(1).expand_as(dz)
    pw = (rois[:, 2] - rois[:, 0] + 1.0).unsqueeze(1).expand_as(dw)
    ph = (rois[:, 3] - rois[:, 1] + 1.0).unsqueeze(1).expand_as(dh)
    pd = (rois[:, 5] - rois[:, 4] + 1.0).unsqueeze(1).expand_as(dd)
    gw, gh, gd = pw * dw.exp(), ph * dh.exp(), pd * dd.exp()
    gx, gy, gz = px + pw * dx, py + ph * dy, pz + pd * dz
    x1, y1 = gx - gw * 0.5 + 0.5, gy - gh * 0.5 + 0.5
    x2, y2 = gx + gw * 0.5 - 0.5, gy + gh * 0.5 - 0.5
    z1, z2 = gz - gd * 0.5 + 0.5, gz + gd * 0.5 - 0.5
    if max_shape is not None:
        x1, x2 = x1.clamp(0, max_shape[1] - 1), x2.clamp(0, max_shape[1] - 1)
        y1, y2 = y1.clamp(0, max_shape[0] - 1), y2.clamp(0, max_shape[0] - 1)
        z1, z2 = z1.clamp(0, max_shape[3] - 1), z2.clamp(0, max_shape[3] - 1)
    return torch.stack([x1, y1, x2, y2, z1, z2], dim=-1).view_as(deltas)


def bbox2delta3d(proposals, gt, means=(0, 0, 0, 0, 0, 0), stds=(1, 1, 1, 1, 1, 1)):
    """Regression targets (dx, dy, dw, dh, dz, dd) of `gt` w.r.t. `proposals`, one kernel.
    Reference: bbox2delta3d, transforms.py:33-63 (about 30 elementwise launches)."""
    assert proposals.size() == gt.size()
    p, g = proposals.float().contiguous(), gt.float().contiguous()
    check_cuda_f32(p, "proposals", ndim=2)
    check_cuda_f32(g, "gt", ndim=2)
    if p.shape[1] < 6:
        raise NotImplementedError("bbox2delta3d: boxes need 6 columns")
    n = p.shape[0]
    out = p.new_empty((n, 6))
    m6 = (ctypes.c_float * 6)(*[float(x) for x in means])
    s6 = (ctypes.c_float * 6)(*[float(x) for x in stds])
    if n:
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib.roi3d_bbox2delta3d(p.data_ptr(), p.shape[1], g.data_ptr(), g.shape[1], n,
                                                   ctypes.addressof(m6), ctypes.addressof(s6), out.data_ptr(),
                                                   stream_ptr()))
    return out
