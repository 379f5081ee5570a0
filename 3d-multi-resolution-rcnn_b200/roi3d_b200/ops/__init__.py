"""Operator surface of the reference's `mmdet.ops` for the 3D RoI hot path (mmdet/ops/__init__.py:1-16):
`nms`, `soft_nms`, `RoIAlign3D`, `roi_align_3d`.  The 2-D ops, RoIPool, DCN and focal loss of that module are
out of scope and are not exported."""
from .nms import nms, nms3d_batched, soft_nms
from .roi_align import RoIAlign3D, RoIAlignFunction3D, roi_align_3d, roi_align_3d_host, set_bug_compat

__all__ = ['nms', 'soft_nms', 'nms3d_batched', 'RoIAlign3D', 'RoIAlignFunction3D', 'roi_align_3d',
           'roi_align_3d_host', 'set_bug_compat']
