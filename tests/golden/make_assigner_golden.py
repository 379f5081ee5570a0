"""Generates tests/golden/assigner_ref.npz: the reference's own MaxIoUAssigner (mmdet/core/bbox/assigners/
max_iou_assigner.py + geometry.py, imported from /root/reference and run on CPU tensors) on seeded boxes, with and
without ignore regions.  Run in the build container (the GPU box has no /root/reference):
    python tests/golden/make_assigner_golden.py
The package __init__ files of the reference import its compiled ops, so the four packages on the path are created as
empty namespace stubs and only the three plain-Python modules are executed."""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402

for name, rel in (("mmdet", "mmdet"), ("mmdet.core", "mmdet/core"), ("mmdet.core.bbox", "mmdet/core/bbox"),
                  ("mmdet.core.bbox.assigners", "mmdet/core/bbox/assigners")):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF, rel)]
    sys.modules[name] = m
MaxIoUAssigner = importlib.import_module("mmdet.core.bbox.assigners.max_iou_assigner").MaxIoUAssigner


def case(n, k, seed):
    rng = np.random.default_rng(seed)
    gt = synth.c1_boxes(max(k, 8), seed=seed)[:k, :6].copy()
    boxes = synth.c1_boxes(n, seed=seed + 1)[:, :6].copy()
    take = rng.integers(0, k, n // 3)
    boxes[: n // 3] = gt[take] + rng.integers(-4, 5, (n // 3, 6)).astype(np.float32)
    boxes[5] = gt[0]
    boxes[6] = gt[0]
    ign = np.concatenate([gt[[1, min(4, k - 1)]], boxes[rng.choice(n, 2, replace=False)] + np.float32(1.0)], 0)
    return boxes.astype(np.float32), gt.astype(np.float32), ign.astype(np.float32)


out = {}
cfgs = []
i = 0
for (n, k, seed) in ((600, 5, 3), (1500, 9, 31)):
    boxes, gt, ign = case(n, k, seed)
    labels = (np.arange(k) % 3 + 1).astype(np.int64)
    for assign_all in (True, False):
        for ignore_thr, wrt in ((-1, True), (0.4, True), (0.4, False)):
            cfg = dict(pos_iou_thr=0.5, neg_iou_thr=0.3, min_pos_iou=0.2, gt_max_assign_all=assign_all,
                       ignore_iof_thr=ignore_thr, ignore_wrt_candidates=wrt)
            res = MaxIoUAssigner(**cfg).assign(torch.from_numpy(boxes), torch.from_numpy(gt),
                                               gt_bboxes_ignore=torch.from_numpy(ign), gt_labels=torch.from_numpy(labels))
            out["boxes_%d" % i], out["gt_%d" % i], out["ign_%d" % i], out["labels_%d" % i] = boxes, gt, ign, labels
            out["cfg_%d" % i] = np.array([0.5, 0.3, 0.2, float(assign_all), ignore_thr, float(wrt)], dtype=np.float64)
            out["gt_inds_%d" % i] = res.gt_inds.numpy()
            out["max_overlaps_%d" % i] = res.max_overlaps.numpy()
            out["assigned_labels_%d" % i] = res.labels.numpy()
            i += 1
out["num_cases"] = np.array(i)
np.savez_compressed(os.path.join(HERE, "assigner_ref.npz"), **out)
print("wrote", i, "cases; positives per case:", [int((out["gt_inds_%d" % j] > 0).sum()) for j in range(i)],
      "ignored (-1) per case:", [int((out["gt_inds_%d" % j] < 0).sum()) for j in range(i)])
