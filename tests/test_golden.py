"""Pins the CPU oracle against fixtures produced by the REFERENCE's own CUDA kernels on a B200
(tests/golden/make_golden.py runs oracle/_ref there; the .npz files are committed).  No GPU, no /root/reference."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    path = os.path.join(HERE, "golden", name)
    if not os.path.exists(path):
        pytest.skip("%s not generated yet" % name)
    return np.load(path)


@pytest.mark.parametrize("tag", ["c7", "m14x10", "b7x3", "a5"])
def test_roi_align_forward_is_bit_identical_to_reference_kernel(oracle, tag):
    z = _load("roi_align3d_ref.npz")
    ps, pdp, sn = (int(v) for v in z["%s_cfg" % tag])
    got = oracle.roi_align3d_forward(z["feats"], z["%s_rois" % tag], ps, pdp, float(z["spatial_scale"]),
                                     float(z["spatial_scale_depth"]), sn, contract=True)
    assert np.array_equal(got, z["%s_out" % tag])
    # the source-literal arithmetic (no FMA contraction) is close but NOT what the compiled kernel computes
    lit = oracle.roi_align3d_forward(z["feats"], z["%s_rois" % tag], ps, pdp, float(z["spatial_scale"]),
                                     float(z["spatial_scale_depth"]), sn, contract=False)
    assert np.abs(lit - z["%s_out" % tag]).max() < 1e-4


@pytest.mark.parametrize("tag", ["c7", "m14x10", "b7x3", "a5"])
def test_roi_align_backward_matches_reference_kernel(oracle, tag):
    """The reference accumulates with fp32 atomics in arbitrary order; the oracle sums in float64.
    Non-cubic outputs: the reference's own top_diff index (bug_compat) is what its kernel executed."""
    z = _load("roi_align3d_ref.npz")
    ps, pdp, sn = (int(v) for v in z["%s_cfg" % tag])
    want = z["%s_gin" % tag]
    got = oracle.roi_align3d_backward(z["%s_gout" % tag], z["%s_rois" % tag], z["feats"].shape,
                                      float(z["spatial_scale"]), float(z["spatial_scale_depth"]), sn,
                                      bug_compat=True, contract=True)
    assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
    if ps != pdp:
        fixed = oracle.roi_align3d_backward(z["%s_gout" % tag], z["%s_rois" % tag], z["feats"].shape,
                                            float(z["spatial_scale"]), float(z["spatial_scale_depth"]), sn,
                                            bug_compat=False, contract=True)
        assert np.abs(fixed - want).max() > 1e-3   # the reference's gradient is wrong here


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_nms_keep_list_equals_reference_kernel(oracle, tag):
    z = _load("nms3d_ref.npz")
    got = oracle.nms3d(z["%s_dets" % tag], float(z["%s_thr" % tag]), contract=True)
    assert np.array_equal(got, z["%s_keep" % tag])


def _assigner_cases():
    path = os.path.join(HERE, "golden", "assigner_ref.npz")
    return range(int(np.load(path)["num_cases"])) if os.path.exists(path) else []


@pytest.mark.parametrize("i", list(_assigner_cases()))
def test_assigner_equals_the_reference_assigner(oracle, i):
    """tests/golden/assigner_ref.npz holds the outputs of the REFERENCE's MaxIoUAssigner (its own Python, CPU tensors;
    tests/golden/make_assigner_golden.py), with and without ignore regions: the oracle restatement gives the same
    gt_inds, max_overlaps and labels, bit for bit."""
    z = _load("assigner_ref.npz")
    pos, neg, mpi, assign_all, ign_thr, wrt = (float(v) for v in z["cfg_%d" % i])
    a, mo, lab = oracle.assign_max_iou(z["boxes_%d" % i], z["gt_%d" % i], z["labels_%d" % i], pos_iou_thr=pos,
                                       neg_iou_thr=neg, min_pos_iou=mpi, gt_max_assign_all=bool(assign_all),
                                       ignore_iof_thr=ign_thr, ignore_wrt_candidates=bool(wrt),
                                       gt_bboxes_ignore=z["ign_%d" % i])
    assert np.array_equal(a, z["gt_inds_%d" % i])
    assert np.array_equal(mo, z["max_overlaps_%d" % i])
    assert np.array_equal(lab, z["assigned_labels_%d" % i])


@pytest.mark.gpu
@pytest.mark.parametrize("i", list(_assigner_cases()))
def test_device_assigner_equals_the_reference_assigner(i):
    """The same fixture against the product (roi3d_assign_max_iou_ignore through the MaxIoUAssigner mirror)."""
    import torch
    from roi3d_b200.core.bbox import MaxIoUAssigner
    z = _load("assigner_ref.npz")
    pos, neg, mpi, assign_all, ign_thr, wrt = (float(v) for v in z["cfg_%d" % i])
    dev = torch.device("cuda:0")
    res = MaxIoUAssigner(pos, neg, mpi, bool(assign_all), ign_thr, bool(wrt)).assign(
        torch.from_numpy(z["boxes_%d" % i]).to(dev), torch.from_numpy(z["gt_%d" % i]).to(dev),
        gt_bboxes_ignore=torch.from_numpy(z["ign_%d" % i]).to(dev), gt_labels=torch.from_numpy(z["labels_%d" % i]).to(dev))
    assert np.array_equal(res.gt_inds.cpu().numpy(), z["gt_inds_%d" % i])
    assert np.array_equal(res.max_overlaps.cpu().numpy(), z["max_overlaps_%d" % i])
    assert np.array_equal(res.labels.cpu().numpy(), z["assigned_labels_%d" % i])
