"""Generates tests/golden/python_ref.npz: outputs of the reference's own pure-Python pieces on the hot path's
neighbouring rows, imported from /root/reference and run on CPU tensors:
  AnchorGenerator3D (gen_base_anchors / grid_anchors / valid_flags)   mmdet/core/anchor/anchor_generator_3d.py
  anchor_inside_flags                                                 mmdet/core/anchor/anchor_target.py:203-217
  bbox2delta3d, delta2bbox3D                                          mmdet/core/bbox/transforms.py:33-63, 105-160
  nms_3d_python                                                       mmdet/core/evaluation/coco_utils.py:245-282
  RandomSampler.sample (numpy RNG, seeded)                            mmdet/core/bbox/samplers/{base,random}_sampler.py
  SingleRoIExtractor.map_roi_levels                                   mmdet/models/roi_extractors/single_level.py:58-76
Run in the build container (the GPU box has no /root/reference):  python tests/golden/make_python_ref_golden.py
Package __init__ files of the reference import compiled ops / absent third-party modules, so packages are empty
namespace stubs, `mmcv` is an empty stub (the functions used here never touch it), and functions that live in modules
with heavier imports are compiled from their own source text (ast), unmodified."""
import ast
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


def stub_pkg(name, rel):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF, rel)]
    sys.modules[name] = m


for name, rel in (("mmdet", "mmdet"), ("mmdet.core", "mmdet/core"), ("mmdet.core.anchor", "mmdet/core/anchor"),
                  ("mmdet.core.bbox", "mmdet/core/bbox"), ("mmdet.core.bbox.assigners", "mmdet/core/bbox/assigners"),
                  ("mmdet.core.bbox.samplers", "mmdet/core/bbox/samplers")):
    stub_pkg(name, rel)
sys.modules["mmcv"] = types.ModuleType("mmcv")


def function_from_source(rel, func, cls=None, extra=None):
    """The reference's own function `func` (of class `cls`) compiled from its source text."""
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    body = tree.body
    if cls is not None:
        body = [n for n in body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    node = [n for n in body if isinstance(n, ast.FunctionDef) and n.name == func][0]
    node.decorator_list = []
    mod = ast.Module(body=[node], type_ignores=[])
    ns = {"torch": torch, "np": np}
    ns.update(extra or {})
    exec(compile(mod, rel, "exec"), ns)
    return ns[func]


out = {}

# ---- anchors -------------------------------------------------------------------------------------------------------
AnchorGenerator3D = importlib.import_module("mmdet.core.anchor.anchor_generator_3d").AnchorGenerator3D
anchor_inside_flags = function_from_source("mmdet/core/anchor/anchor_target.py", "anchor_inside_flags")
ANCHOR_CASES = [
    ((5, 8, 6), 8, 4, [2], [2], [1.0], (5, 7, 6), (56, 48, 3, 20), 0),
    ((7, 9, 11), 4, 2, [2, 4], [2, 3], [0.5, 1.0, 2.0], (6, 9, 10), (36, 42, 3, 13), 3),
    ((3, 4, 5), 16, 8, [8], [2], [1.0], (3, 4, 5), (64, 80, 3, 24), -1),
]
import inspect  # noqa: E402
print("AnchorGenerator3D.__init__", inspect.signature(AnchorGenerator3D.__init__))
print("grid_anchors", inspect.signature(AnchorGenerator3D.grid_anchors), "valid_flags", inspect.signature(AnchorGenerator3D.valid_flags))
for i, (fm, st, sd, scales, dscales, ratios, valid, img_shape, border) in enumerate(ANCHOR_CASES):
    gen = AnchorGenerator3D(st, scales, dscales, ratios, sd)
    a = gen.grid_anchors(fm, st, sd, device='cpu')
    v = gen.valid_flags(fm, valid, device='cpu')
    f = anchor_inside_flags(a, v, img_shape, border)
    out["anchor_base_%d" % i] = gen.base_anchors.numpy()
    out["anchor_grid_%d" % i] = a.numpy()
    out["anchor_valid_%d" % i] = v.numpy().astype(np.uint8)
    out["anchor_inside_%d" % i] = f.numpy().astype(np.uint8)
out["anchor_cases"] = np.array(len(ANCHOR_CASES))

# ---- bbox2delta3d / delta2bbox3D -------------------------------------------------------------------------------------
tr = importlib.import_module("mmdet.core.bbox.transforms")
rng = np.random.default_rng(0)
lo, sz = rng.uniform(0, 400, (2000, 3)).astype(np.float32), rng.uniform(4, 60, (2000, 3)).astype(np.float32)
p = np.stack([lo[:, 0], lo[:, 1], lo[:, 0] + sz[:, 0], lo[:, 1] + sz[:, 1], lo[:, 2], lo[:, 2] + sz[:, 2]], 1).astype(np.float32)
g = (p + rng.uniform(-1.5, 1.5, p.shape)).astype(np.float32)
means, stds = [0.0] * 6, [0.1, 0.1, 0.2, 0.2, 0.1, 0.2]
d = tr.bbox2delta3d(torch.from_numpy(p), torch.from_numpy(g), means, stds)
out["t_props"], out["t_gt"], out["t_stds"] = p, g, np.array(stds, np.float32)
out["t_deltas"] = d.numpy()
dd = (rng.standard_normal((2000, 6)) * 0.5).astype(np.float32)
print("delta2bbox3D", inspect.signature(tr.delta2bbox3D))
back = tr.delta2bbox3D(torch.from_numpy(p), torch.from_numpy(dd), means, stds, max_shape=(384, 420, 3, 96))
out["t_rand_deltas"], out["t_decoded"] = dd, back.numpy()
out["t_max_shape"] = np.array([384, 420, 3, 96])

# ---- evaluation-time NMS (numpy) --------------------------------------------------------------------------------------
nms_3d_python = function_from_source("mmdet/core/evaluation/coco_utils.py", "nms_3d_python")
for i, (n, seed) in enumerate(((300, 3), (77, 9))):
    dets = synth.c1_boxes(n, seed=seed)
    dets[:, :6] = np.round(dets[:, :6])           # json boxes are integers in practice; keep some exact ties
    if n > 50:
        dets[10, :6] = dets[3, :6]
    keep = nms_3d_python(np.arange(n), dets.copy(), 0.1)   # json_results = the indices: returns json_results[keep]
    out["e_dets_%d" % i] = dets
    out["e_keep_%d" % i] = np.asarray(keep, dtype=np.int64)
out["e_cases"] = np.array(2)

# ---- RandomSampler ----------------------------------------------------------------------------------------------------
RandomSampler = importlib.import_module("mmdet.core.bbox.samplers.random_sampler").RandomSampler
AssignResult = importlib.import_module("mmdet.core.bbox.assigners.assign_result").AssignResult
for i, cfg in enumerate((dict(num=512, pos_fraction=0.25, neg_pos_ub=-1, add_gt_as_proposals=True),
                         dict(num=256, pos_fraction=0.5, neg_pos_ub=3, add_gt_as_proposals=False))):
    r2 = np.random.default_rng(40 + i)
    n, k = 3000, 7
    gi = np.zeros(n, np.int64)
    gi[r2.choice(n, 900, replace=False)] = r2.integers(1, k + 1, 900)
    gi[r2.choice(n, 200, replace=False)] = -1
    boxes = synth.c1_boxes(n, seed=50 + i)[:, :6]
    gtb = synth.c1_boxes(k + 8, seed=60 + i)[:k, :6]
    labels = (np.arange(k) % 3 + 1).astype(np.int64)
    ar = AssignResult(k, torch.from_numpy(gi.copy()), torch.zeros(n), labels=torch.from_numpy(labels[np.maximum(gi, 1) - 1] * (gi > 0)))
    np.random.seed(1234 + i)
    res = RandomSampler(**cfg).sample(ar, torch.from_numpy(boxes.copy()), torch.from_numpy(gtb), torch.from_numpy(labels))
    out["s_gt_inds_%d" % i], out["s_boxes_%d" % i], out["s_gt_%d" % i], out["s_labels_%d" % i] = gi, boxes, gtb, labels
    out["s_cfg_%d" % i] = np.array([cfg["num"], cfg["pos_fraction"], cfg["neg_pos_ub"], float(cfg["add_gt_as_proposals"])])
    out["s_seed_%d" % i] = np.array(1234 + i)
    out["s_pos_inds_%d" % i], out["s_neg_inds_%d" % i] = res.pos_inds.numpy(), res.neg_inds.numpy()
    out["s_pos_bboxes_%d" % i], out["s_pos_assigned_%d" % i] = res.pos_bboxes.numpy(), res.pos_assigned_gt_inds.numpy()
out["s_cases"] = np.array(2)

# ---- map_roi_levels ----------------------------------------------------------------------------------------------------
map_roi_levels = function_from_source("mmdet/models/roi_extractors/single_level.py", "map_roi_levels", cls="SingleRoIExtractor")


class _Self(object):
    finest_scale = 56


rois = np.concatenate([synth.c2_rois(400, seed=8), synth.c2_rois(100, seed=9, img=(2048, 2048, 160))], 0).astype(np.float32)
# exact boundaries of the level rule: sqrt(w * h) = 56 * 2^j
for j, s in enumerate((112.0, 224.0, 448.0)):
    rois[j] = [0, 10, 10, 10 + s - 1, 10 + s - 1, 3, 9]
lv = map_roi_levels(_Self(), torch.from_numpy(rois), 4)
out["m_rois"], out["m_levels"] = rois, lv.numpy()
np.savez_compressed(os.path.join(HERE, "python_ref.npz"), **out)
print("wrote python_ref.npz:", {k: out[k].shape for k in ("anchor_grid_1", "t_deltas", "e_keep_0", "s_pos_inds_0", "m_levels")},
      "levels", np.bincount(out["m_levels"], minlength=4))
