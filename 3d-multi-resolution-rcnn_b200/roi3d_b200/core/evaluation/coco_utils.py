"""Evaluation-time 3D NMS on the device.

The reference filters the json detections of every volume once more before COCO-style scoring:
`apply_nms` (mmdet/core/evaluation/coco_utils.py:306-332) loops over the volumes, collects each volume's results
with a Python scan over ALL json results, and runs `nms_3d_python` (:245-282), a numpy float64 greedy NMS at
iou 0.1.  Here every volume is one segment of ONE batched device NMS (`roi3d_nms3d_eval_batched`) whose IoU is
evaluated in float64 in numpy's operation order, so the kept sets and their order are the reference's.

Same call signatures as the reference.  The `filter_based_on_precomputed_proposals` branch reads a pickle that is not
part of the repository (coco_utils.py:307-311) and is not reproduced.
"""
import numpy as np
import torch

from ... import _lib
from ..._util import stream_ptr, workspace


def nms_3d_eval_batched(boxes_per_volume, iou_thr, device=None):
    """boxes_per_volume: list of [n_i, 7] arrays (x1,y1,x2,y2,z1,z2,score).  Returns a list of int64 index arrays,
    each in the reference's return order (descending score; equal scores: lower index first)."""
    nseg = len(boxes_per_volume)
    counts = [len(b) for b in boxes_per_volume]
    n_max = max(counts) if counts else 0
    if nseg == 0 or n_max == 0:
        return [np.zeros(0, dtype=np.int64) for _ in range(nseg)]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    host = np.zeros((nseg, n_max, 7), dtype=np.float32)
    for s, b in enumerate(boxes_per_volume):
        if counts[s]:
            host[s, :counts[s]] = np.asarray(b, dtype=np.float32).reshape(-1, 7)
    dets = torch.from_numpy(host).to(dev)
    seg = torch.tensor(counts, dtype=torch.int32, device=dev)
    keep = torch.empty((nseg, n_max), dtype=torch.int64, device=dev)
    keep_s = torch.empty((nseg, n_max), dtype=torch.int64, device=dev)
    num = torch.zeros((nseg,), dtype=torch.int32, device=dev)
    nbytes = _lib.lib.roi3d_nms3d_workspace_bytes(nseg, n_max)
    _buf, ws = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.roi3d_nms3d_eval_batched(
            dets.data_ptr(), seg.data_ptr(), nseg, n_max, float(iou_thr), keep.data_ptr(), keep_s.data_ptr(),
            num.data_ptr(), ws, nbytes, stream_ptr()))
    num_h = num.cpu().numpy()
    keep_h = keep_s.cpu().numpy()
    return [keep_h[s, :num_h[s]].copy() for s in range(nseg)]


def nms_3d_python(json_results, boxes, iou_thr):
    """Drop-in for coco_utils.py:245-282: returns the kept `json_results` (numpy object array) in score order."""
    if len(boxes) == 0:
        return []
    keep = nms_3d_eval_batched([np.asarray(boxes, dtype=np.float64)], iou_thr)[0]
    return np.array(json_results)[keep]


def apply_nms(full_filename_to_id, json_results, nms_thresh=0.1, score_thresh=0,
              filter_based_on_precomputed_proposals=False):
    """Drop-in for coco_utils.py:306-332.  One pass groups the results by image id (the reference rescans the whole
    list per volume), one batched device NMS filters every volume."""
    if filter_based_on_precomputed_proposals:
        raise NotImplementedError("the precomputed-proposal filter reads a pickle that is not part of the reference "
                                  "repository (coco_utils.py:307-311)")
    by_id = {}
    for r in json_results:
        by_id.setdefault(r['image_id'], []).append(r)
    ids = [img_id for _fn, img_id in full_filename_to_id.items()]
    groups = [by_id.get(i, []) for i in ids]
    boxes = [np.asarray([r['original_bbox'] for r in g], dtype=np.float64).reshape(-1, 7) for g in groups]
    kept = nms_3d_eval_batched(boxes, nms_thresh)
    out = []
    for g, k in zip(groups, kept):
        for i in k:
            if g[i]['score'] < score_thresh:
                continue
            out.append(g[i])
    return out
