"""Dense 3D IoU on the device.  Reference: bbox_overlaps, 6-column non-aligned branch, mmdet/core/bbox/geometry.py:49-60
(about 25 torch elementwise launches over broadcast [m, n] temporaries); here one kernel."""
import torch

from ... import _lib
from ..._util import check_cuda_f32, stream_ptr


def bbox_overlaps(bboxes1, bboxes2, mode='iou', is_aligned=False):
    """bboxes1 [m, >=6], bboxes2 [n, >=6] fp32 CUDA (x1,y1,x2,y2,z1,z2,...) -> ious [m, n]."""
    assert mode in ['iou', 'iof']
    # mode='iof': the reference's 6-column branch never looks at `mode` (geometry.py:49-60 divide by the union whatever
    # it says), so 'iof' returns the IoU here as it does there
    if is_aligned:
        raise NotImplementedError("is_aligned=True is not on the 3D path (the reference's aligned branch stops in a "
                                  "debugger, geometry.py:33)")
    check_cuda_f32(bboxes1, "bboxes1", ndim=2)
    check_cuda_f32(bboxes2, "bboxes2", ndim=2)
    if bboxes1.shape[1] < 6 or bboxes2.shape[1] < 6:
        raise NotImplementedError("bbox_overlaps: only 3D boxes (>= 6 columns) are supported")
    m, n = bboxes1.shape[0], bboxes2.shape[0]
    out = bboxes1.new_empty((m, n))
    if m * n == 0:
        return out
    b1, b2 = bboxes1.contiguous(), bboxes2.contiguous()
    with torch.cuda.device(b1.device):
        _lib.check(_lib.lib.roi3d_bbox_overlaps3d(b1.data_ptr(), m, b1.shape[1], b2.data_ptr(), n, b2.shape[1],
                                                  out.data_ptr(), stream_ptr()))
    return out
