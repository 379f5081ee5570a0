"""Mask paste of the 3D mask head on the device.

Reference: FCNMaskHead3D.get_seg_masks, mmdet/models/mask_heads/fcn_mask_head_3d.py:124-187 -- for every detection,
on the host: sigmoid -> numpy, `skimage.transform.resize` to the box size, threshold, paste into a full-volume uint8
array.  Here every detection's resize + threshold runs in one launch of `roi3d_mask_paste` (csrc/mask_paste.cu);
`paste_masks_compact` returns the box-sized binary masks (what a B200 pipeline keeps), `get_seg_masks` returns the
reference's structure (a full-volume array per detection, grouped by class).
"""
import numpy as np
import torch

from ... import _lib
from ..._util import check_cuda_f32, stream_ptr


def paste_masks_compact(mask_pred, det_bboxes, det_labels, mask_thr_binary, scale_factor=1.0, class_agnostic=False):
    """mask_pred [n, num_classes, Dm, Hm, Wm] CUDA fp32 LOGITS (the head's output before sigmoid, as get_seg_masks
    receives it); det_bboxes [n, >=6]; det_labels [n] (0-based foreground labels).
    Returns (boxes int32 [n, 6] numpy, masks: list of uint8 numpy arrays of shape (d, h, w), labels numpy 1-based)."""
    check_cuda_f32(mask_pred, "mask_pred", ndim=5)
    n = mask_pred.shape[0]
    bboxes = det_bboxes.detach().cpu().numpy()[:, :6]
    labels = det_labels.detach().cpu().numpy() + 1
    # the reference's integer box: (bboxes[i, :] / scale_factor).astype(np.int32)   (fcn_mask_head_3d.py:163)
    boxes = (bboxes / scale_factor).astype(np.int32).reshape(-1, 6)
    if n == 0:
        return boxes, [], labels
    w = np.maximum(boxes[:, 2] - boxes[:, 0] + 1, 1).astype(np.int64)
    h = np.maximum(boxes[:, 3] - boxes[:, 1] + 1, 1).astype(np.int64)
    d = np.maximum(boxes[:, 5] - boxes[:, 4] + 1, 1).astype(np.int64)
    sizes = w * h * d
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    dev = mask_pred.device
    ch = torch.zeros(n, dtype=torch.long, device=dev) if class_agnostic else torch.from_numpy(labels).to(dev).long()
    sel = mask_pred.detach()[torch.arange(n, device=dev), ch].contiguous()            # [n, Dm, Hm, Wm]
    boxes_dev = torch.from_numpy(boxes).to(dev)
    off_dev = torch.from_numpy(offsets).to(dev)
    out = torch.empty(int(sizes.sum()), dtype=torch.uint8, device=dev)
    Dm, Hm, Wm = sel.shape[1:]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.roi3d_mask_paste(sel.data_ptr(), n, Dm, Hm, Wm, boxes_dev.data_ptr(), off_dev.data_ptr(),
                                             float(mask_thr_binary), out.data_ptr(), stream_ptr()))
    flat = out.cpu().numpy()
    masks = [flat[offsets[i]:offsets[i] + sizes[i]].reshape(d[i], h[i], w[i]) for i in range(n)]
    return boxes, masks, labels


def get_seg_masks(mask_pred, det_bboxes, det_labels, rcnn_test_cfg, ori_shape, scale_factor, rescale, num_classes,
                  class_agnostic=False):
    """Drop-in for FCNMaskHead3D.get_seg_masks (same arguments plus the head's num_classes / class_agnostic):
    returns cls_segms, a list per foreground class of full-volume uint8 masks (img_d, img_h, img_w)."""
    if not rescale:
        raise NotImplementedError("rescale=False stops in a debugger in the reference (fcn_mask_head_3d.py:153-157)")
    thr = rcnn_test_cfg['mask_thr_binary'] if isinstance(rcnn_test_cfg, dict) else rcnn_test_cfg.mask_thr_binary
    img_h, img_w, img_d = ori_shape[:3]
    boxes, masks, labels = paste_masks_compact(mask_pred, det_bboxes, det_labels, thr, scale_factor, class_agnostic)
    cls_segms = [[] for _ in range(num_classes - 1)]
    for i, m in enumerate(masks):
        b = boxes[i]
        dd, hh, ww = m.shape
        im_mask = np.zeros((img_d, img_h, img_w), dtype=np.uint8)
        im_mask[b[4]:b[4] + dd, b[1]:b[1] + hh, b[0]:b[0] + ww] = m
        cls_segms[labels[i] - 1].append(im_mask)
    return cls_segms
