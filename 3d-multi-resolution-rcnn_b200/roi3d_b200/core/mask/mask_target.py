"""Mask targets of the 3D mask head on the device (training half).

Reference: mask_target / mask_target_single, mmdet/core/mask/mask_target.py:8-50 -- for every positive proposal, on
the host: `.cpu().numpy()`, crop of the assigned ground-truth mask to the proposal's int32 box,
`255 * skimage.transform.resize(crop, (mask_size_depth, mask_size, mask_size))`, `.astype(np.uint8)`, non-zero -> 1,
stack, back to the device.  Here the crops are never copied to the host: one launch of `roi3d_mask_target`
(csrc/mask_paste.cu) resamples every proposal's crop; only the [n, 6] proposals are read on the host (they size the
workspace), as the reference reads them too.
"""
import numpy as np
import torch

from ... import _lib
from ..._util import stream_ptr, workspace


def _cfg_get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def mask_target_single(pos_proposals, pos_assigned_gt_inds, gt_masks, cfg):
    """pos_proposals [n, 6] CUDA fp32; pos_assigned_gt_inds [n] int64; gt_masks [G, D, H, W] uint8 (CUDA, or a host
    tensor / array that is uploaded once); cfg.mask_size, cfg.mask_size_depth.  Returns [n, Md, Ms, Ms] fp32 of 0/1."""
    ms, md = int(_cfg_get(cfg, 'mask_size')), int(_cfg_get(cfg, 'mask_size_depth'))
    n = pos_proposals.size(0)
    dev = pos_proposals.device
    if n == 0:
        return pos_proposals.new_zeros((0, ms, ms))  # the reference's (2-D shaped) empty result, mask_target.py:48
    if pos_proposals.shape[1] != 6:
        raise NotImplementedError("mask_target_single: only 3D proposals (x1,y1,x2,y2,z1,z2) are on this path")
    if not dev.type == 'cuda':
        raise NotImplementedError("mask_target_single: the B200 path has no CPU implementation")
    gm = torch.as_tensor(gt_masks)
    if gm.dtype != torch.uint8:
        gm = gm.to(torch.uint8)
    gm = gm.to(dev).contiguous()
    G, D, H, W = gm.shape
    boxes = np.ascontiguousarray(pos_proposals.detach().cpu().numpy().astype(np.int32))   # mask_target.py:25,29
    inds = np.ascontiguousarray(pos_assigned_gt_inds.detach().cpu().numpy().astype(np.int64))
    w = np.maximum(boxes[:, 2] - boxes[:, 0] + 1, 1)
    h = np.maximum(boxes[:, 3] - boxes[:, 1] + 1, 1)
    d = np.maximum(boxes[:, 5] - boxes[:, 4] + 1, 1)
    crops = np.ascontiguousarray(np.stack([np.minimum(boxes[:, 4] + d, D) - boxes[:, 4],
                                           np.minimum(boxes[:, 1] + h, H) - boxes[:, 1],
                                           np.minimum(boxes[:, 0] + w, W) - boxes[:, 0]], 1).astype(np.int32))
    nbytes = _lib.lib.roi3d_mask_target_workspace_bytes(crops.ctypes.data, n)
    _buf, ws = workspace(dev, nbytes)
    out = torch.empty((n, md, ms, ms), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.roi3d_mask_target(gm.data_ptr(), G, D, H, W, boxes.ctypes.data, inds.ctypes.data, n, md, ms, ms,
                                              out.data_ptr(), ws, nbytes, stream_ptr()))
    return out


def mask_target(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list, cfg):
    """mask_target, mask_target.py:8-14: per image, concatenated."""
    return torch.cat([mask_target_single(p, i, g, cfg)
                      for p, i, g in zip(pos_proposals_list, pos_assigned_gt_inds_list, gt_masks_list)])
