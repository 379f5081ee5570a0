"""MaxIoUAssigner over the fused device kernel.

Reference: mmdet/core/bbox/assigners/max_iou_assigner.py:8-171.  Same constructor arguments and `assign` signature;
the [gts x boxes] IoU matrix, its two reductions and the Python loop over the gts are one call of
`roi3d_assign_max_iou` (csrc/assign.cu) -- the matrix is never stored, nothing syncs the host.
"""
import torch

from .... import _lib
from ...._util import check_cuda_f32, stream_ptr, workspace
from .assign_result import AssignResult


class MaxIoUAssigner(object):
    """-1 = don't care, 0 = negative, i > 0 = positive for gt i (1-based); see the reference docstring (:9-32)."""

    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True, ignore_iof_thr=-1,
                 ignore_wrt_candidates=True):
        self.pos_iou_thr = pos_iou_thr
        self.neg_iou_thr = neg_iou_thr
        self.min_pos_iou = min_pos_iou
        self.gt_max_assign_all = gt_max_assign_all
        self.ignore_iof_thr = ignore_iof_thr
        self.ignore_wrt_candidates = ignore_wrt_candidates

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        if isinstance(gt_bboxes, list) and len(gt_bboxes) == 1:
            gt_bboxes = gt_bboxes[0]
        if isinstance(gt_labels, list) and len(gt_labels) == 1:
            gt_labels = gt_labels[0]
        if bboxes.shape[0] == 0 or gt_bboxes.shape[0] == 0:
            raise ValueError('No gt or bboxes')
        use_ignore = (self.ignore_iof_thr > 0) and (gt_bboxes_ignore is not None) and (gt_bboxes_ignore.numel() > 0)
        check_cuda_f32(bboxes, "bboxes", ndim=2)
        check_cuda_f32(gt_bboxes, "gt_bboxes", ndim=2)
        if bboxes.shape[1] < 6 or gt_bboxes.shape[1] < 6:
            raise NotImplementedError("MaxIoUAssigner: only 3D boxes (>= 6 columns) are supported")
        if gt_bboxes.shape[1] != 6:  # the kernel reads gt rows at a stride of 6 floats (geometry.py:49 takes 6 columns)
            gt_bboxes = gt_bboxes[:, :6]
        if isinstance(self.neg_iou_thr, tuple):
            assert len(self.neg_iou_thr) == 2
            neg_lo, neg_hi = float(self.neg_iou_thr[0]), float(self.neg_iou_thr[1])
        elif isinstance(self.neg_iou_thr, float):
            neg_lo, neg_hi = 0.0, float(self.neg_iou_thr)
        else:  # the reference assigns no negatives for any other type (max_iou_assigner.py:146-152)
            neg_lo, neg_hi = 1.0, 0.0
        b, g = bboxes.detach().contiguous(), gt_bboxes.detach().contiguous()
        n, k, dev = b.shape[0], g.shape[0], b.device
        gt_inds = torch.empty((n,), dtype=torch.long, device=dev)
        max_overlaps = torch.empty((n,), dtype=torch.float32, device=dev)
        labels = None
        lab_ptr = gl_ptr = None
        if gt_labels is not None:
            gl = gt_labels.to(device=dev, dtype=torch.long).contiguous()
            labels = torch.empty((n,), dtype=torch.long, device=dev)
            lab_ptr, gl_ptr = labels.data_ptr(), gl.data_ptr()
        ign = None
        if use_ignore:
            # max_iou_assigner.py:101-111.  For 6-column boxes the reference's bbox_overlaps never looks at mode='iof'
            # (geometry.py:49-60): the "iof" against the ignore regions is the plain IoU, in either orientation.
            from ..geometry import bbox_overlaps
            check_cuda_f32(gt_bboxes_ignore, "gt_bboxes_ignore", ndim=2)
            gi = gt_bboxes_ignore.detach()[:, :6].contiguous()
            b6 = b[:, :6]
            if self.ignore_wrt_candidates:
                ign_max = bbox_overlaps(b6, gi, mode='iof').max(dim=1)[0]
            else:
                ign_max = bbox_overlaps(gi, b6, mode='iof').max(dim=0)[0]
            ign = (ign_max > self.ignore_iof_thr).to(torch.uint8).contiguous()
        nbytes = _lib.lib.roi3d_assign_workspace_bytes(n, k)
        _buf, ws = workspace(dev, nbytes)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.roi3d_assign_max_iou_ignore(
                b.data_ptr(), n, b.shape[1], g.data_ptr(), k, gl_ptr, None if ign is None else ign.data_ptr(),
                float(self.pos_iou_thr), neg_lo, neg_hi, float(self.min_pos_iou), 1 if self.gt_max_assign_all else 0,
                gt_inds.data_ptr(), max_overlaps.data_ptr(), lab_ptr, ws, nbytes, stream_ptr()))
        return AssignResult(k, gt_inds, max_overlaps, labels=labels)
