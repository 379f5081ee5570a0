"""CPU tests of the oracle itself: the reference's known-answer vectors, internal consistency (forward/backward
adjointness, contraction modes, 3D NMS against the reference's numpy restatement), the glue restatements."""
import numpy as np
import pytest

import synth


# mmdet/core/bbox/geometry.py:81-102 (bbox_overlaps_test): the only golden vectors the reference holds
IOU_KAT = [
    ([2, 3, 4, 6, 3, 4], [2, 3, 4, 6, 3, 4], 1.0),
    ([39, 63, 203, 112, 4, 5], [54, 66, 198, 114, 4, 5], 0.798),
    ([49, 75, 203, 125, 4, 5], [42, 78, 186, 126, 4, 5], 0.7899),
    ([31, 69, 201, 125, 4, 5], [18, 63, 235, 135, 4, 5], 0.6125),
]


@pytest.mark.parametrize("contract", [True, False])
def test_iou_known_answers(oracle, contract):
    for a, b, gold in IOU_KAT:
        assert round(oracle.iou3d(a, b, contract), 4) == gold
        assert round(oracle.iou3d(b, a, contract), 4) == gold


def test_nms_matches_reference_numpy_nms(oracle):
    """oracle.nms3d (CUDA semantics) against the reference's own CPU 3D NMS, nms_3d_python
    (coco_utils.py:245-282), on integer boxes where fp32 and fp64 IoU agree exactly."""
    rng = np.random.default_rng(3)
    n = 600
    c = rng.integers(0, 100, (n, 3))
    s = rng.integers(4, 30, (n, 3))
    boxes = np.stack([c[:, 0], c[:, 1], c[:, 0] + s[:, 0], c[:, 1] + s[:, 1], c[:, 2], c[:, 2] + s[:, 2]], 1)
    scores = rng.permutation(np.linspace(0.1, 0.9, n))
    dets = np.concatenate([boxes, scores[:, None]], 1).astype(np.float32)
    for thr in (0.1, 0.3, 0.5, 0.7):
        keep, by_score = oracle.nms3d(dets, thr, return_score_order=True)
        ref = oracle.nms_3d_python(dets.astype(np.float64), thr)
        assert np.array_equal(by_score, ref)
        assert np.array_equal(keep, np.sort(ref))


def test_nms_edge_cases(oracle):
    assert oracle.nms3d(np.zeros((0, 7), np.float32), 0.5).shape == (0,)
    one = np.array([[0, 0, 5, 5, 0, 5, 0.3]], np.float32)
    assert oracle.nms3d(one, 0.5).tolist() == [0]
    # equal scores: stable order = lower index wins
    two = np.array([[0, 0, 9, 9, 0, 9, 0.5], [0, 0, 9, 9, 0, 9, 0.5]], np.float32)
    assert oracle.nms3d(two, 0.5).tolist() == [0]
    # disjoint in z only: 3D semantics keeps both (the reference's CPU wrapper would not)
    z = np.array([[0, 0, 10, 10, 0, 5, .9], [0, 0, 10, 10, 100, 105, .8]], np.float32)
    assert oracle.nms3d(z, 0.5).tolist() == [0, 1]
    assert oracle.nms_cpu_2d(z, 0.5).tolist() == [1]


def test_mask_matches_greedy(oracle):
    dets = synth.c1_boxes(300, seed=5)
    order = oracle.argsort_desc_stable(dets[:, 6])
    mask = oracle.nms3d_mask(dets[order], 0.7)
    n, cb = mask.shape
    remv = np.zeros(cb, dtype=np.uint64)
    kept = []
    for i in range(n):
        if not (int(remv[i // 64]) >> (i % 64)) & 1:
            kept.append(i)
            remv |= mask[i]
    assert np.array_equal(np.sort(order[kept]), oracle.nms3d(dets, 0.7))


def test_roi_align_forward_backward_adjoint(oracle):
    rng = np.random.default_rng(0)
    f = rng.standard_normal((2, 6, 9, 14, 15)).astype(np.float32)
    rois = np.concatenate([synth.adversarial_rois((9, 14, 15), 0.25, 0.5, batch=2),
                           synth.c2_rois(6, seed=1, img=(60, 56, 18), batch=2)], 0)
    for (ps, pd, sn) in [(7, 7, 2), (7, 3, 2), (4, 5, 0), (14, 10, 1)]:
        ok = (rois[:, 3] >= rois[:, 1]) & (rois[:, 4] >= rois[:, 2]) & (rois[:, 6] >= rois[:, 5])
        rr = rois[ok] if sn == 0 else rois   # adaptive sampling of a zero-size RoI is 0/0 = NaN (as in the reference)
        out = oracle.roi_align3d_forward(f, rr, ps, pd, 0.25, 0.5, sn)
        g = rng.standard_normal(out.shape).astype(np.float32)
        gi = oracle.roi_align3d_backward(g, rr, f.shape, 0.25, 0.5, sn)
        lhs = float((out.astype(np.float64) * g).sum())
        rhs = float((f.astype(np.float64) * gi).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs)), (ps, pd, sn, lhs, rhs)


def test_roi_align_constant_map_and_outside(oracle):
    f = np.full((1, 3, 8, 12, 12), 2.5, np.float32)
    rois = np.array([[0, 4, 4, 30, 30, 2, 10], [0, -100, -100, -60, -60, -50, -30]], np.float32)
    out = oracle.roi_align3d_forward(f, rois, 7, 7, 0.25, 0.5, 2)
    assert np.allclose(out[0], 2.5, atol=1e-6)
    assert np.all(out[1] == 0)


def test_roi_align_contract_modes_agree_to_rounding(oracle):
    rng = np.random.default_rng(1)
    f = rng.standard_normal((1, 4, 10, 20, 20)).astype(np.float32)
    rois = synth.c2_rois(16, seed=3, img=(80, 80, 20))
    a = oracle.roi_align3d_forward(f, rois, 7, 7, 0.25, 0.5, 2, contract=True)
    b = oracle.roi_align3d_forward(f, rois, 7, 7, 0.25, 0.5, 2, contract=False)
    assert np.abs(a - b).max() < 2e-5


def test_bug_compat_only_differs_for_noncubic(oracle):
    rng = np.random.default_rng(2)
    rois = synth.c2_rois(4, seed=5, img=(60, 60, 18))
    shape = (1, 3, 9, 15, 15)
    g = rng.standard_normal((4, 3, 7, 7, 7)).astype(np.float32)
    a = oracle.roi_align3d_backward(g, rois, shape, 0.25, 0.5, 2, bug_compat=False)
    b = oracle.roi_align3d_backward(g, rois, shape, 0.25, 0.5, 2, bug_compat=True)
    assert np.array_equal(a, b)
    g = rng.standard_normal((4, 3, 3, 7, 7)).astype(np.float32)
    a = oracle.roi_align3d_backward(g, rois, shape, 0.25, 0.5, 2, bug_compat=False)
    b = oracle.roi_align3d_backward(g, rois, shape, 0.25, 0.5, 2, bug_compat=True)
    assert not np.array_equal(a, b)


def test_map_roi_levels(oracle):
    def roi(w, h, d):
        return [0, 0, 0, w - 1, h - 1, 0, d - 1]
    rois = np.array([roi(10, 10, 10), roi(56, 56, 1), roi(30, 30, 14), roi(112, 112, 1), roi(60, 60, 60),
                     roi(500, 500, 100), roi(1, 1, 1)], np.float32)
    lv = oracle.map_roi_levels(rois, 4)
    # sqrt(w*h*d): 31.6, 56, 112.2, 112, 464.8, 5000, 1
    assert lv.tolist() == [0, 0, 1, 1, 3, 3, 0]
    assert oracle.map_roi_levels(rois, 2).max() == 1


def test_delta2bbox_identity_and_clamp(oracle):
    anchors = np.array([[10, 20, 29, 49, 4, 11], [100, 100, 131, 131, 30, 45]], np.float32)
    z = np.zeros((2, 6), np.float32)
    out = oracle.delta2bbox3d(anchors, z, max_shape=(512, 512, 3, 160))
    assert np.allclose(out, anchors, atol=1e-4)
    big = np.full((2, 6), 100.0, np.float32)
    out = oracle.delta2bbox3d(anchors, big, max_shape=(512, 512, 3, 160))
    assert out[:, [0, 2]].max() <= 511 and out[:, [1, 3]].max() <= 511 and out[:, [4, 5]].max() <= 159
    assert out.min() >= 0


def test_topk_tie_rule(oracle):
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1], np.float32)
    assert oracle.topk(s, 3).tolist() == [1, 3, 0]
    assert oracle.topk(s, 10).tolist() == [1, 3, 0, 2, 4]


def test_grid_anchor_order_matches_permute(oracle):
    """Anchor i must pair with score i of permute(2,3,1,0).reshape(-1) (rpn_head_3d.py:87-94)."""
    D, H, W = 3, 4, 5
    base = oracle.gen_base_anchors(4, [2], [2], [1.0], 2)
    assert base.shape == (1, 6)
    anchors = oracle.grid_anchors(base, (D, H, W), 4, 2)
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing='ij')
    cx = np.transpose((xx * 4.0)[None], (2, 3, 1, 0)).reshape(-1)
    cz = np.transpose((zz * 2.0)[None], (2, 3, 1, 0)).reshape(-1)
    assert np.allclose((anchors[:, 0] + anchors[:, 2]) / 2 - (base[0, 0] + base[0, 2]) / 2, cx)
    assert np.allclose((anchors[:, 4] + anchors[:, 5]) / 2 - (base[0, 4] + base[0, 5]) / 2, cz)


def test_get_bboxes_single_smoke(oracle):
    rng = np.random.default_rng(6)
    dims = [(8, 16, 16), (4, 8, 8)]
    strides, dstrides = [4, 8], [2, 4]
    cls = [2 * rng.standard_normal((1,) + d).astype(np.float32) for d in dims]
    reg = [0.1 * rng.standard_normal((6,) + d).astype(np.float32) for d in dims]
    anchors = [oracle.grid_anchors(oracle.gen_base_anchors(s, [2], [2], [1.0], ds), d, s, ds)
               for d, s, ds in zip(dims, strides, dstrides)]
    props = oracle.get_bboxes_single(cls, reg, anchors, (64, 64, 3, 16), nms_pre=200, nms_post=100, max_num=120,
                                     nms_thr=0.7)
    assert props.shape[1] == 7 and 0 < props.shape[0] <= 120
    assert np.all(np.diff(props[:, 6]) <= 0)
    assert props[:, :6].min() >= 0 and props[:, 2].max() <= 63 and props[:, 5].max() <= 15


def test_multiclass_nms(oracle):
    rng = np.random.default_rng(8)
    d = synth.c1_boxes(200, seed=9)
    scores = np.stack([1 - d[:, 6], d[:, 6]], 1).astype(np.float32)
    bb, lab = oracle.multiclass_nms_3d(d[:, :6], scores, 0.2, 0.5, max_num=50)
    assert bb.shape[0] <= 50 and bb.shape[1] == 7 and np.all(lab == 0)
    assert np.all(bb[:, 6] > 0.2)
    bb, lab = oracle.multiclass_nms_3d(d[:, :6], scores, 2.0, 0.5, max_num=50)
    assert bb.shape == (0, 7) and lab.shape == (0,)


def test_apply_nms_restatement_groups_by_volume_and_keeps_score_order(oracle):
    """oracle.apply_nms (coco_utils.py:306-332): per volume in dict order, survivors in descending score, score
    threshold applied after NMS; results of other volumes never interact."""
    box = [0, 0, 9, 9, 0, 9]
    res = [dict(image_id=2, original_bbox=box + [0.6], score=0.6, tag="b-low"),
           dict(image_id=1, original_bbox=box + [0.9], score=0.9, tag="a-high"),
           dict(image_id=2, original_bbox=box + [0.8], score=0.8, tag="b-high"),
           dict(image_id=1, original_bbox=[50, 50, 59, 59, 20, 29, 0.2], score=0.2, tag="a-far-weak"),
           dict(image_id=1, original_bbox=[1, 0, 10, 9, 0, 9, 0.7], score=0.7, tag="a-overlap")]
    out = oracle.apply_nms({"a": 1, "b": 2, "c": 3}, res, 0.1, 0.3)
    assert [r["tag"] for r in out] == ["a-high", "b-high"]
    out = oracle.apply_nms({"b": 2, "a": 1}, res, 1.0, 0.0)  # iou <= 1 keeps everything: pure ordering
    assert [r["tag"] for r in out] == ["b-high", "b-low", "a-high", "a-overlap", "a-far-weak"]


def test_bbox_overlaps3d_known_answers_of_the_reference(oracle):
    """bbox_overlaps_test, mmdet/core/bbox/geometry.py:81-102: the reference's own known answers (rounded to 4 dp)
    and the 2 x 3 shape case."""
    cases = [([39, 63, 203, 112, 4, 5], [54, 66, 198, 114, 4, 5], 0.798),
             ([49, 75, 203, 125, 4, 5], [42, 78, 186, 126, 4, 5], 0.7899),
             ([31, 69, 201, 125, 4, 5], [18, 63, 235, 135, 4, 5], 0.6125),
             ([2, 3, 4, 6, 3, 4], [2, 3, 4, 6, 3, 4], 1.0)]
    for a, b, want in cases:
        v = oracle.bbox_overlaps3d(np.array([a], np.float32), np.array([b], np.float32))
        assert v.shape == (1, 1) and round(float(v[0, 0]), 4) == want
    r = oracle.bbox_overlaps3d(np.array([[2, 3, 4, 6, 3, 4], [39, 63, 203, 112, 4, 5]], np.float32),
                               np.array([[2, 3, 4, 6, 3, 4], [54, 66, 198, 114, 4, 5], [49, 75, 203, 125, 4, 5]],
                                        np.float32))
    assert r.shape == (2, 3) and int(r[0, 0]) == 1


def test_assign_max_iou_rules_in_order(oracle):
    """max_iou_assigner.py:128-171: -1 default, negatives, positives, then every gt claims its best boxes (later gts
    override earlier ones), labels follow."""
    gt = np.array([[0, 0, 9, 9, 0, 9], [0, 0, 9, 9, 0, 9], [50, 50, 59, 59, 20, 29]], np.float32)  # gt 0 == gt 1
    boxes = np.array([[0, 0, 9, 9, 0, 9],          # IoU 1 with gts 0 and 1: argmax -> gt 0, rule 4 -> gt 1 (later wins)
                      [2, 0, 11, 9, 0, 9],         # 0.667 with gts 0/1: between neg 0.3 and pos 0.7 -> stays -1
                      [200, 200, 209, 209, 40, 49],  # no overlap -> negative
                      [54, 50, 63, 59, 20, 29],    # 0.43 with gt 2 only, but it is gt 2's best box -> positive via rule 4
                      [8, 0, 17, 9, 0, 9]], np.float32)  # 0.11 -> negative
    a, mo, lab = oracle.assign_max_iou(boxes, gt, [5, 6, 7], 0.7, 0.3, 0.3, True)
    assert a.tolist() == [2, -1, 0, 3, 0] and lab.tolist() == [6, 0, 0, 7, 0]
    assert mo[0] == 1.0 and mo[2] == 0.0
    a2, _, _ = oracle.assign_max_iou(boxes, gt, None, 0.7, (0.05, 0.3), 0.5, False)  # tuple band; min_pos_iou blocks gt 2
    assert a2.tolist() == [2, -1, -1, -1, 0]


def test_bbox2delta3d_inverts_delta2bbox3d(oracle):
    rng = np.random.default_rng(0)
    lo = rng.uniform(0, 100, (50, 3)).astype(np.float32)
    sz = rng.uniform(4, 60, (50, 3)).astype(np.float32)
    p = np.stack([lo[:, 0], lo[:, 1], lo[:, 0] + sz[:, 0], lo[:, 1] + sz[:, 1], lo[:, 2], lo[:, 2] + sz[:, 2]], 1)
    g = p + rng.uniform(-3, 3, p.shape).astype(np.float32)
    means, stds = (0.0,) * 6, (0.1, 0.1, 0.2, 0.2, 0.1, 0.2)
    d = oracle.bbox2delta3d(p, g, means, stds)
    back = oracle.delta2bbox3d(p, d, means, stds, None)
    assert np.abs(back - g).max() < 1e-3
